// C ABI of aiolos_b200 (include/aiolos_b200.h): thin exception-to-error-code layer over aio::Engine.
#include "aiolos_b200.h"
#include "engine.h"
#include "hooks.h"
#include "cuda_compat.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

struct aio_ensemble {
    std::unique_ptr<aio::Engine> eng;
    std::vector<aio::Problem> problems;      // host-side description of every column (kept for the writers)
    std::vector<std::vector<double>> monitor_prev;   // previous_monitor_J / _T per column (io.cpp:1048-1054)
};

namespace aio { void full2_selftest(int ncols, int N2, int Bout, int S, const double* coef, double* x); }   // rad_full2.cu
namespace aio { void measure_device_peaks(int device, double* fp64_tflops, double* copy_gbs); }   // peaks.cu

namespace {
thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

template <typename F> int guarded(F&& f) {
    try { return f(); }
    catch (const std::invalid_argument& e) { return fail(AIO_ERR_ARG, e.what()); }
    catch (const std::runtime_error& e) {
        const std::string m = e.what();
        if (m.find("unsupported configuration") != std::string::npos) return fail(AIO_ERR_UNSUPPORTED, m);
        if (m.find("CUDA") != std::string::npos) return fail(AIO_ERR_CUDA, m);
        if (m.find("cannot open") != std::string::npos || m.find("cannot read") != std::string::npos) return fail(AIO_ERR_IO, m);
        return fail(AIO_ERR_INTERNAL, m);
    }
    catch (const std::exception& e) { return fail(AIO_ERR_INTERNAL, e.what()); }
    catch (...) { return fail(AIO_ERR_INTERNAL, "unknown exception"); }
}

// Per-column overrides on a problem that came from a dump (no .par file to re-read): the few keys whose effect on the
// constructor's result is local are re-applied here with the constructor's own expressions.
void apply_dump_override(aio::Problem& q, const std::string& key, const std::string& value) {
    const double v = std::stod(value);
    const double pi = 3.141592653589793, au = 1.495978707e13;
    if (key == "PARI_TIME_TMAX") { q.t_max = v; return; }
    if (key == "PARI_UVSTAR" || key == "PARI_XSTAR") {
        if (q.solar_heating_thermal.size() != (size_t)q.num_bands_in) throw std::runtime_error("dump lacks solar_heating_thermal");
        if (q.fluxfile_used) throw std::runtime_error("unsupported configuration: PARI_UVSTAR override with FLUX_FILE");
        if (key == "PARI_UVSTAR") q.UV_star = v; else q.X_star = v;
        for (int b = 0; b < q.num_bands_in; b++) {          // init_and_bounds.cpp:808-846
            double F = q.solar_heating_thermal[b];
            if (q.num_bands_in == 1) {
                F += q.UV_star / (4. * pi * pow(q.planet_semimajor * au, 2.));
                F += q.X_star / (4. * pi * pow(q.planet_semimajor * au, 2.));
            } else if (q.l_i_in[b + 1] <= 0.09161) {
                if (q.l_i_in[b + 1] <= 0.0161) F += q.X_star / (4. * pi * pow(q.planet_semimajor * au, 2.));
                else F += q.UV_star / (4. * pi * pow(q.planet_semimajor * au, 2.));
            }
            q.solar_heating_final[b] = F;
            if (b < (int)q.solar_heating.size()) q.solar_heating[b] = F;
        }
        return;
    }
    throw std::invalid_argument("override key '" + key + "' is not supported on a dump-created ensemble");
}
}  // namespace

extern "C" {

const char* aio_last_error(void) { return g_err.c_str(); }
const char* aio_version(void) {
#ifdef AIO_EMU_BUILD
    static const std::string base = "aiolos_b200 0.1 (CTA emulator, test build)";
#else
    static const std::string base = "aiolos_b200 0.1 (sm_100a)";
#endif
    // a library made with `make PROBLEM=<file>` names the problem plugin it carries (hooks.h)
    static const std::string full = aio::hooks_compiled() ? base + " problem=" + aio::hooks_name() : base;
    return full.c_str();
}

int aio_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int aio_selftest_block_thomas(int ncols, int num_cells2, int bands_out, int num_species, const double* coef, double* x) {
    if (!coef || !x) return fail(AIO_ERR_ARG, "aio_selftest_block_thomas: null pointer");
    return guarded([&]() { aio::full2_selftest(ncols, num_cells2, bands_out, num_species, coef, x); return AIO_OK; });
}

int aio_device_peaks(int device, double* fp64_tflops, double* copy_gbs) {
#ifdef AIO_EMU_BUILD
    (void)device; (void)fp64_tflops; (void)copy_gbs;
    return fail(AIO_ERR_CUDA, "aio_device_peaks: CUDA device required (emulator build)");
#else
    return guarded([&]() { aio::measure_device_peaks(device, fp64_tflops, copy_gbs); return AIO_OK; });
#endif
}

int aio_create(const char* dir, const char* par, const char* spc, int ncols, const aio_override* ov, int n_ov, int device,
               aio_ensemble** out) {
    if (!dir || !par || !out || ncols < 1 || (n_ov > 0 && !ov)) return fail(AIO_ERR_ARG, "aio_create: bad argument");
    *out = nullptr;
    return guarded([&]() {
        std::unique_ptr<aio_ensemble> e(new aio_ensemble);
        e->problems.resize(ncols);
        // columns without a private override share one parse + init
        aio::Problem base; bool have_base = false;
        for (int c = 0; c < ncols; c++) {
            std::vector<std::pair<std::string, std::string>> mine;
            bool priv = false;
            for (int k = 0; k < n_ov; k++) {
                if (!ov[k].key || !ov[k].value) throw std::invalid_argument("aio_create: null override");
                if (ov[k].column < 0 || ov[k].column == c) { mine.emplace_back(ov[k].key, ov[k].value); if (ov[k].column == c) priv = true; }
            }
            std::string err;
            if (!priv && have_base) { e->problems[c] = base; continue; }
            if (!aio::problem_from_files(dir, par, spc ? spc : "", mine, e->problems[c], err)) throw std::runtime_error(err);
            if (!priv) { base = e->problems[c]; have_base = true; }
        }
        e->eng.reset(new aio::Engine(e->problems, device));
        *out = e.release();
        return AIO_OK;
    });
}

int aio_create_from_dump(const char* path, int ncols, const aio_override* ov, int n_ov, int device, aio_ensemble** out) {
    if (!path || !out || ncols < 1 || (n_ov > 0 && !ov)) return fail(AIO_ERR_ARG, "aio_create_from_dump: bad argument");
    *out = nullptr;
    return guarded([&]() {
        aio::DumpMap d; std::string err;
        if (!aio::read_dump(path, d, err)) throw std::runtime_error(err);
        aio::Problem base;
        if (!aio::problem_from_dump(d, base, err)) throw std::runtime_error(err);
        std::unique_ptr<aio_ensemble> e(new aio_ensemble);
        e->problems.assign(ncols, base);
        for (int k = 0; k < n_ov; k++) {
            if (!ov[k].key || !ov[k].value) throw std::invalid_argument("aio_create_from_dump: null override");
            if (ov[k].column >= ncols) throw std::invalid_argument("aio_create_from_dump: override column out of range");
            for (int c = 0; c < ncols; c++)
                if (ov[k].column < 0 || ov[k].column == c) apply_dump_override(e->problems[c], ov[k].key, ov[k].value);
        }
        e->eng.reset(new aio::Engine(e->problems, device));
        *out = e.release();
        return AIO_OK;
    });
}

void aio_destroy(aio_ensemble* e) { delete e; }

int aio_dims(aio_ensemble* e, int* ncols, int* num_cells, int* num_species, int* bands_in, int* bands_out) {
    if (!e) return fail(AIO_ERR_ARG, "null handle");
    if (ncols) *ncols = e->eng->ncol();
    if (num_cells) *num_cells = e->eng->num_cells();
    if (num_species) *num_species = e->eng->num_species();
    if (bands_in) *bands_in = e->eng->num_bands_in();
    if (bands_out) *bands_out = e->eng->num_bands_out();
    return AIO_OK;
}

int aio_advance(aio_ensemble* e, long long max_steps, double* device_ms) {
    if (!e || max_steps < 0) return fail(AIO_ERR_ARG, "aio_advance: bad argument");
    return guarded([&]() { double ms = e->eng->advance(max_steps, device_ms != nullptr); if (device_ms) *device_ms = ms; return AIO_OK; });
}

int aio_run(aio_ensemble* e, double t_stop, long long max_steps, long long* steps_done) {
    if (!e) return fail(AIO_ERR_ARG, "null handle");
    return guarded([&]() {
        if (std::isfinite(t_stop)) e->eng->clamp_tmax(t_stop);
        long long n = e->eng->run_to_tmax(256, max_steps > 0 ? max_steps : (1LL << 62));
        if (steps_done) *steps_done = n;
        return AIO_OK;
    });
}

int aio_kernel_times(aio_ensemble* e, long long nsteps, double* ms3) {
    if (!e || !ms3 || nsteps < 1) return fail(AIO_ERR_ARG, "aio_kernel_times: bad argument");
    return guarded([&]() { e->eng->kernel_times(nsteps, ms3); return AIO_OK; });
}

int aio_set_tmax(aio_ensemble* e, int column, double t_max) {
    if (!e) return fail(AIO_ERR_ARG, "null handle");
    return guarded([&]() { e->eng->set_tmax(column, t_max); return AIO_OK; });
}

int aio_set_option(aio_ensemble* e, const char* key, const char* value) {
    if (!e || !key || !value) return fail(AIO_ERR_ARG, "aio_set_option: bad argument");
    return guarded([&]() { e->eng->set_option(key, value); return AIO_OK; });
}

int aio_status(aio_ensemble* e, int column, double* t, double* dt, long long* steps, int* state) {
    if (!e || column < 0 || column >= e->eng->ncol()) return fail(AIO_ERR_ARG, "aio_status: bad argument");
    return guarded([&]() { e->eng->status(column, t, dt, steps, state); return AIO_OK; });
}

int aio_status_all(aio_ensemble* e, double* t, double* dt, long long* steps, int* state) {
    if (!e) return fail(AIO_ERR_ARG, "null handle");
    return guarded([&]() { e->eng->status_all(t, dt, steps, state); return AIO_OK; });
}

int aio_get(aio_ensemble* e, int column, int species, int field, double* dst, int n) {
    if (!e || !dst || n < 0) return fail(AIO_ERR_ARG, "aio_get: bad argument");
    return guarded([&]() { e->eng->get_field(column, species, field, dst, n); return AIO_OK; });
}

int aio_problem_field(aio_ensemble* e, int column, const char* name, double* dst, int capacity) {
    if (!e || !name || column < 0 || column >= (int)e->problems.size() || capacity < 0) { fail(AIO_ERR_ARG, "aio_problem_field: bad argument"); return AIO_ERR_ARG; }
    std::vector<double> v;
    if (!aio::problem_field(e->problems[column], name, v)) return fail(AIO_ERR_ARG, std::string("aio_problem_field: unknown field ") + name);
    if (dst) for (int k = 0; k < capacity && k < (int)v.size(); k++) dst[k] = v[k];
    return (int)v.size();
}

int aio_upload_state(aio_ensemble* e, const double* u) {
    if (!e || !u) return fail(AIO_ERR_ARG, "aio_upload_state: bad argument");
    return guarded([&]() { e->eng->upload_u(u); return AIO_OK; });
}
int aio_upload_state_async(aio_ensemble* e, const double* u) {
    if (!e || !u) return fail(AIO_ERR_ARG, "aio_upload_state_async: bad argument");
    return guarded([&]() { e->eng->upload_u_async(u); return AIO_OK; });
}
int aio_commit_upload(aio_ensemble* e) {
    if (!e) return fail(AIO_ERR_ARG, "null handle");
    return guarded([&]() { e->eng->commit_upload(); return AIO_OK; });
}
int aio_download_state(aio_ensemble* e, double* u) {
    if (!e || !u) return fail(AIO_ERR_ARG, "aio_download_state: bad argument");
    return guarded([&]() { e->eng->download_u(u); return AIO_OK; });
}
int aio_column_mass(aio_ensemble* e, double* m) {
    if (!e || !m) return fail(AIO_ERR_ARG, "aio_column_mass: bad argument");
    return guarded([&]() { e->eng->column_mass(m); return AIO_OK; });
}

int aio_write_outputs(aio_ensemble* e, int column, int index, const char* outdir) {
    if (!e || column < 0 || column >= e->eng->ncol()) return fail(AIO_ERR_ARG, "aio_write_outputs: bad argument");
    return guarded([&]() { aio::write_outputs(*e->eng, e->problems[column], column, index, outdir ? outdir : ""); return AIO_OK; });
}

int aio_write_monitor(aio_ensemble* e, int column, const char* outdir, int truncate) {
    if (!e || column < 0 || column >= e->eng->ncol()) return fail(AIO_ERR_ARG, "aio_write_monitor: bad argument");
    return guarded([&]() {
        if (e->monitor_prev.size() != (size_t)e->eng->ncol()) e->monitor_prev.assign(e->eng->ncol(), {});
        if (truncate) e->monitor_prev[column].clear();
        aio::write_monitor(*e->eng, e->problems[column], column, outdir ? outdir : "", truncate != 0, e->monitor_prev[column]);
        return AIO_OK;
    });
}

long long aio_kernel_launches(aio_ensemble* e) { return e ? e->eng->kernel_launches() : 0; }
int aio_kernel_name(aio_ensemble* e, int slot, char* dst, int capacity) {
    if (!e || !dst || capacity < 1) return fail(AIO_ERR_ARG, "aio_kernel_name: bad argument");
    return guarded([&]() {
        const std::string s = e->eng->kernel_name(slot);
        snprintf(dst, (size_t)capacity, "%s", s.c_str());
        return AIO_OK;
    });
}
long long aio_total_steps(aio_ensemble* e) {
    if (!e) return 0;
    try { return e->eng->total_steps(); } catch (...) { return -1; }
}

}  // extern "C"
