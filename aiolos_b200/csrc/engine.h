// Host-side handle of an ensemble of independent columns resident on one GPU (see engine.cu).
#pragma once
#include "problem.h"
#include <stdexcept>
#include <string>
#include <vector>

namespace aio {

class Engine {
public:
    Engine(const std::vector<Problem>& columns, int device);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    double advance(long long nsteps, bool timed);              // returns device milliseconds when timed
    long long run_to_tmax(long long chunk, long long max_total);
    void set_tmax(int col, double t_max);                       // col < 0: all columns
    void clamp_tmax(double t_stop);                             // t_max = min(t_max, t_stop) for every column
    void column_mass(double* mass_per_column);                  // sum_s sum_j u1*vol, reduced on the device
    long long total_steps();
    void kernel_times(long long nsteps, double* out3);          // avg device ms per step of {hydro, friction, radiation}
    void set_option(const std::string& key, const std::string& value);
    void status(int col, double* t, double* dt, long long* steps, int* state);
    void status_all(double* t, double* dt, long long* steps, int* state);
    double col_scalar(int col, int which);
    void get_field(int col, int species, int field, double* dst, int n);
    void upload_u(const double* src);
    void upload_u_async(const double* src);
    void commit_upload();
    void download_u(double* dst);
    void sync();
    int ncol() const; int num_cells() const; int num_species() const; int num_bands_out() const; int num_bands_in() const;
    long long kernel_launches() const;
    std::string kernel_name(int slot) const;                    // what aio_kernel_times slot 0/1/2 launches
    void* stream() const;
    const Problem& problem0() const { return prob0; }

private:
    struct Impl;
    Impl* im;
    Problem prob0;
    void construct(const std::vector<Problem>& columns, int device);
    void destroy();
    void launch_step();
    void launch_fused(long long nsteps);
    void launch_friction();
    void launch_full();
    void full_setup();
    void fetch_states();
    void push_states();
};

// Per-species output / diagnostic files in the reference's column layout (io.cpp:704-932).
void write_outputs(Engine& eng, const Problem& prob, int col, int index, const std::string& outdir);
void write_monitor(Engine& eng, const Problem& prob, int col, const std::string& outdir, bool truncate, std::vector<double>& prev);

}  // namespace aio
