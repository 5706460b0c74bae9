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
    else if (p.S <= 16) AIO_LAUNCH(k_friction<16>, ncol, NT, 0, im->stream, p);
    else throw std::runtime_error("unsupported configuration: more than 16 species with friction");
}

// Full FLD (PARI_USE_RADIATION 1): rad_full2.cuh (register-tiled LU, warp per column) when the block fits 24 unknowns, else the
// first-generation kernel of rad_full.cuh (one warp per column, LU in shared memory).
void Engine::full_setup() {
    DevParams& p = im->p;
    p.Xstore = nullptr; p.fcoef = nullptr; p.Tsol = nullptr; p.dcore = nullptr; p.ssurf = nullptr;
    if (p.use_rad_fluxes != 1) return;
    const FullDims fd = full_dims(p.Bout, p.S, p.Bin);
    if (fd.n > 32 || p.Bin > 32) throw std::runtime_error("unsupported configuration: full FLD with more than 32 unknowns per cell");
    const Full2Dims f2 = full2_dims(p.Bout, p.S);
    im->full2_ok = fd.n <= 24;
    size_t words = (size_t)p.ncol * p.N2 * fd.n * (p.Bout + 1);
    size_t extra = 0;
    if (im->full2_ok) {
        words = std::max(words, (size_t)p.ncol * p.N2 * f2.XB);
        extra = (size_t)p.ncol * p.N2 * f2.NC + (size_t)p.ncol * p.S * p.NP + (size_t)p.ncol * (p.Bout + p.Bin);
    }
    size_t free_b = 0, total_b = 0;
#ifndef AIO_EMU_BUILD
    CK(cudaMemGetInfo(&free_b, &total_b));
    if ((words + extra) * sizeof(double) > free_b * 9 / 10)
        throw std::runtime_error("unsupported configuration: full FLD factor storage does not fit in HBM; use fewer columns per GPU");
#endif
    p.Xstore = im->dalloc<double>(words);
    im->nt_full = 64;
    im->smem_full = (size_t)(im->nt_full / 32) * fd.per_warp * sizeof(double);
    int maxsm = 0; int dev = 0;
#ifndef AIO_EMU_BUILD
    CK(cudaGetDevice(&dev));
#endif
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (im->smem_full > (size_t)maxsm) throw std::runtime_error("unsupported configuration: full FLD block does not fit in shared memory");
    CK(cudaFuncSetAttribute(k_rad_full<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)im->smem_full));
    CK(cudaFuncSetAttribute(k_rad_full<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)im->smem_full));
    CK(cudaFuncSetAttribute(k_rad_full<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)im->smem_full));
    CK(cudaFuncSetAttribute(k_rad_full<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)im->smem_full));
    CK(cudaFuncSetAttribute(k_rad_full<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)im->smem_full));
    if (im->full2_ok) {
        p.fcoef = im->dalloc<double>((size_t)p.ncol * p.N2 * f2.NC);
        p.Tsol = im->dalloc<double>((size_t)p.ncol * p.S * p.NP);
        p.dcore = im->dalloc<double>((size_t)p.ncol * p.Bout);
        p.ssurf = im->dalloc<double>((size_t)p.ncol * p.Bin);
        im->full_ri = f2.RI;
        im->smem_full2 = (size_t)f2.sm_warp * sizeof(double);
        // stellar heating by k_full_heat2 when its [Bin][N2+1] shared array leaves room for two CTAs per SM
        im->smem_heat2 = ((size_t)p.Bin * ((p.N2 + 1) | 1) + p.Bin) * sizeof(double);
        if (im->smem_heat2 > (size_t)maxsm / 2) im->smem_heat2 = 0;
        full2_configure(im->full_ri, im->smem_full2, im->smem_heat2);
        im->full_mode = 2;
    }
}
void Engine::launch_full() {
    DevParams& p = im->p;
    if (im->full_mode == 2) {
        full2_launch(p, im->full_ri, im->smem_full2, im->smem_heat2, (void*)im->stream);
        im->launches += 3;
        return;
    }
    const int wpb = im->nt_full / 32;
    const int n = p.Bout + p.S, grid = (p.ncol + wpb - 1) / wpb;
    if (im->full_mode == 0) AIO_LAUNCH(k_rad_full<0>, grid, im->nt_full, im->smem_full, im->stream, p);       // option full_lu=smem
    else if (n <= 8) AIO_LAUNCH(k_rad_full<8>, grid, im->nt_full, im->smem_full, im->stream, p);
    else if (n <= 16) AIO_LAUNCH(k_rad_full<16>, grid, im->nt_full, im->smem_full, im->stream, p);
    else if (n <= 24) AIO_LAUNCH(k_rad_full<24>, grid, im->nt_full, im->smem_full, im->stream, p);
    else AIO_LAUNCH(k_rad_full<32>, grid, im->nt_full, im->smem_full, im->stream, p);
}

// Name of the kernel(s) the engine launches for one slot of aio_kernel_times (0 hydro / fused step, 1 friction, 2 radiation),
// from the same switches launch_step() dispatches on.
std::string Engine::kernel_name(int slot) const {
    const DevParams& p = im->p;
    const std::string cpt = std::to_string(im->CPT);
    if (slot == 0) {
        if (im->fused && im->v2) return "k_step1<" + cpt + "," + std::to_string(im->NT) + "," + std::to_string(im->step1_ctas) + ">";
        if (im->v2) return "k_hydro2<" + cpt + ">";
        return "k_hydro<" + cpt + (im->NT > 256 ? "," + std::to_string(im->NT) : "") + ">";
    }
    if (slot == 1) {
        if (!im->has_friction) return "";
        const int S = p.S;
        return "k_friction<" + std::to_string(S <= 2 ? 2 : S <= 3 ? 3 : S <= 4 ? 4 : S <= 6 ? 6 : S <= 8 ? 8 : 16) + ">";
    }
    if (slot == 2) {
        std::string pre = p.kap_cell ? "k_opacity + " : "";
        if (p.use_rad_fluxes == 1) {
            if (im->full_mode == 2) return pre + (im->smem_heat2 ? "k_full_heat2" : "k_full_heat") + " + k_full_asm2 + k_full_lu<" + std::to_string(im->full_ri) + "> + k_full_post";
            return pre + "k_rad_full<" + std::to_string(im->full_mode == 0 ? 0 : (p.Bout + p.S + 7) / 8 * 8) + ">";
        }
        if (p.photochemistry_level == 1) return pre + "k_photochem";
        if (p.use_rad_fluxes == 2) {
            if (im->fused && im->v2) return "";
            if (im->v2 && p.S == 1 && !p.Tsave && !p.kap_cell) return "k_rad3<" + cpt + ">";
            if (im->v2 && !(p.use_collisional_heating && p.S > 1) && !p.kap_cell) return "k_rad2<" + cpt + ">";
            return pre + "k_rad_simple<" + cpt + ">";
        }
        return "";
    }
    throw std::invalid_argument("kernel_name: slot outside 0..2");
}

}  // namespace aio
