// Launch helpers that dispatch on the species count (included at the end of engine.cu).
namespace aio {

void Engine::launch_friction() {
    DevParams& p = im->p;
    const int ncol = p.ncol;
    const int NT = (p.N2 <= 32) ? 32 : (p.N2 <= 64) ? 64 : 128;
    if (p.S <= 2) AIO_LAUNCH(k_friction<2>, ncol, NT, 0, im->stream, p);
    else if (p.S <= 3) AIO_LAUNCH(k_friction<3>, ncol, NT, 0, im->stream, p);
    else if (p.S <= 4) AIO_LAUNCH(k_friction<4>, ncol, NT, 0, im->stream, p);
    else if (p.S <= 6) AIO_LAUNCH(k_friction<6>, ncol, NT, 0, im->stream, p);
    else if (p.S <= 8) AIO_LAUNCH(k_friction<8>, ncol, NT, 0, im->stream, p);
    else throw std::runtime_error("unsupported configuration: more than 8 species with friction");
}

void Engine::full_setup() {}
void Engine::launch_full() { throw std::runtime_error("unsupported configuration: PARI_USE_RADIATION 1 kernel not built yet"); }

}  // namespace aio
