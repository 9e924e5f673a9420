// FBS feeder tables shared by the kernels (gp_fbs.cu) and the specialised-kernel generator
// (gp_fbs_jit.cu).
#pragma once
#include "gp_common.cuh"

#include <string>
#include <vector>

namespace gp {

struct __align__(16) FbsStep {
    int bus_vrow[3];
    int kind;
    int par_vrow[3];
    int child_begin;
    int edge_irow[3];
    int child_count;
    int load_srow[3];
    int pad0;
    double cap[3];
    double pad1;
    double gamma[3];
    double pad2;
    double inv_gamma[3];
    double pad3;
    double2 z[9];
    int gam_row[3];   // row of the per-scenario gamma array (GpSolveExtras.gamma_dev) per regulated phase, else -1
    int pad4;
};
static_assert(sizeof(FbsStep) == 64 + 32 + 32 + 32 + 144 + 16, "FbsStep layout");

// Destinations of the fused all-gather (GpSolveExtras.peer_vmag / peer_loss), passed to kernels by value
struct PeerDst {
    double* vmag[GP_MAX_PEERS];
    double2* loss[GP_MAX_PEERS];
    int n;
    int pad;
};

struct FbsConst {
    int root_vrow[3];
    int root_load_srow[3];
    double root_cap[3];
    double2 vref[3];
    double aPQ, aI, aZ;
    // volt-var droop evaluated inside the iteration (GpSolveExtras.vv_*, set per launch): breakpoints V1..V4, the
    // two ramps as slope / intercept (computed on the host the way VoltVARController.get_Q does), minQ, maxQ
    double vv_v[4], vv_lo[2], vv_hi[2], vv_q[2];
};

struct FbsImpl {
    int n_steps = 0, n_vrows = 0, n_irows = 0, n_srows = 0, n_child = 0, n_lines = 0;
    int n_grows = 0;              // regulated (step, phase) pairs = rows of a per-scenario gamma array
    FbsStep* d_steps = nullptr;
    int4* d_child = nullptr;  // [n_child] {I rows of the child edge, 0}
    // the streaming kernel's own copy of the two tables with shared I rows (see gp_fbs_create): an I row
    // that only repeats a child's current reads that child's row and is not stored during the sweeps, a
    // row that stays zero is dropped; d_alias lists {row, source row or -1} for the one copy-out at the end
    FbsStep* d_steps_a = nullptr;
    int4* d_child_a = nullptr;
    int2* d_alias = nullptr;
    int n_alias = 0;
    int* d_lines = nullptr;   // [n_lines][9] irow, tx vrow, rx vrow
    int* d_vrow_srow = nullptr;   // [n_vrows] spu row of each V row or -1
    double* d_vrow_cap = nullptr; // [n_vrows]
    FbsConst k;
    // fused depth-first schedule (fbs_dfs_kernel): 2 events per step, or none if not applicable
    int2* d_events = nullptr;
    int4* d_evrows = nullptr;   // per event: the 3 rows to prefetch, .w = 1 for I rows (descend), 0 for spu rows
    int n_events = 0, depth = 0;
    bool dfs_stack_ok = true;     // fbs_dfs_kernel's extra conditions (regulators cover their bus, depth <= 64)
    // on-chip kernel (fbs_onchip_kernel): tile scheduler words {next tile, warps done}, SM count,
    // warps per block that fit the feeder's V and I rows in shared memory (0 = does not fit)
    void* d_walk = nullptr;       // event table of fbs_walk_kernel (WalkEv[n_walk]); n_walk = 0: the feeder is not walked
    int n_walk = 0;
    int* d_unfed = nullptr;       // V rows no forward step writes (kept at the flat start's zero)
    int n_unfed = 0;
    // One pair of self-resetting scheduler words PER STREAM (sched_words()): the words are only correct for
    // one kernel in flight, launches on one stream are ordered, launches on different streams (the chunked
    // host path, callers with their own streams) each get their own pair.
    static constexpr int NSCHED = 32;
    unsigned* d_sched = nullptr;                  // [NSCHED + 1][4] words; pair NSCHED is the overflow pair
    cudaStream_t sched_stream[NSCHED] = {};
    int n_sched_streams = 0;
    int n_sm = 0, onchip_warps = 0;
    // feeder-specialised on-chip kernel (gp_fbs_specialise): host copies of the tables the
    // generator reads, the loaded module and its kernel
    std::vector<FbsStep> h_steps;
    std::vector<int4> h_child;
    std::vector<int> h_lines;     // [n_lines][9] irow, tx vrow, rx vrow
    void* jit_module = nullptr;   // CUmodule
    void* jit_kernel = nullptr;   // CUfunction
    void* jit_module_host = nullptr;  // the read-ahead build of the one-launch host path (built on first use)
    void* jit_kernel_host = nullptr;
    bool jit_host_tried = false;
    void* jit_module_ex = nullptr;    // the build with per-scenario regulator ratios and warm starts (built on first use)
    void* jit_kernel_ex = nullptr;
    bool jit_ex_tried = false;
    int jit_warps = 0;
    int jit_warps_host = 0;       // warps per block of the host path's build (gp_set_option("host_warps"))
    bool jit_ig = false;          // the specialised kernel keeps the I rows in global memory (the I array)
    bool jit_dfs = false, jit_plan_dfs = false;   // the depth-first build (V rows in the caller's array, I slots + a stack on chip)
    int jit_alt_warps = 0;        // warps and I placement of the two-sweep plan (the host path's build when jit_dfs)
    bool jit_alt_ig = false;
    int n_dfs_stack = -1;         // its stack slots (-1: the feeder cannot be built that way)
    int jit_rows = 0;             // shared-memory rows per warp of the built kernel
    int n_islots = 0;             // shared-memory slots the specialised kernel needs for the I rows (fbs_jit_islots)
    int jit_plan_warps = 0;       // warps per block gp_fbs_specialise will use (0 = feeder does not fit)
    bool jit_plan_ig = false;
    void* d_zc = nullptr;         // device staging rows of the one-launch host path
    size_t zc_bytes = 0;
    // host staging for gp_fbs_solve_batch_host
    static constexpr int NSTREAM = 4;
    cudaStream_t streams[NSTREAM] = {nullptr, nullptr, nullptr, nullptr};
    void* d_stage[NSTREAM] = {nullptr, nullptr, nullptr, nullptr};
    size_t stage_bytes = 0;
};


// gp_fbs_jit.cu: CUDA source of the feeder-specialised on-chip kernel, NVRTC compilation
std::string fbs_jit_source(const std::vector<FbsStep>& steps, const std::vector<int4>& child,
                           const std::vector<int>& lines, const FbsConst& k, int n_vrows, int n_irows, int n_srows,
                           int nw, bool ig, bool read_ahead = false, bool ex = false, bool dfs = false);
int fbs_jit_dfs_stack(const std::vector<FbsStep>& steps, const std::vector<int4>& child, const FbsConst& k, int n_vrows,
                      int n_irows);
int fbs_jit_islots(const std::vector<FbsStep>& steps, const std::vector<int4>& child, int n_irows,
                   std::vector<int>& slot, std::vector<char>& owner);
int jit_compile(const std::string& src, std::vector<char>& cubin, std::string& log);

}  // namespace gp
