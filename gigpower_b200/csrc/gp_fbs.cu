// Batched Forward-Backward Sweep for sm_100a.
//
// One thread owns one scenario and walks the feeder tree exactly in the
// reference's order (solution_fbs.py:87-128): forward sweep over topo_order,
// ZIP loads, backward sweep over reversed(topo_order) including the parent
// voltage rewrite, root-mismatch convergence test.  Scenarios are independent,
// so there is no inter-thread communication; a thread leaves the loop at the
// iteration its scenario converges, which is the per-scenario "freeze" that
// 1e-9 parity needs.
//
// Memory: state lives in HBM as packed rows [row][ld] of complex128 with the
// scenario index innermost, so every access of a warp is one 512-byte coalesced
// 128-bit transaction.  The per-feeder step table is scenario-invariant; all
// lanes read the same words (broadcast through L1/the read-only path).
#include "gp_fbs.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <functional>
#include <new>
#include <vector>

namespace gp {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
extern int g_nr3_prefetch;     // gp_nr3.cu
extern int g_nr3_kernel;
extern int g_nr3_ring;
}  // namespace gp
extern int g_peer_push_streams;   // gp_peer.cu
namespace gp {
int nr3_gamma_rows(GpFeeder* f);   // gp_nr3.cu
static int g_jit_dfs = -1;         // gp_set_option("jit_dfs", -1|0|1): depth-first build of the specialised kernel (V rows off chip)
static int g_jit_i_global = -1;    // gp_set_option("jit_i_global", -1|0|1): I rows of the specialised kernel in global memory
static int g_host_read_ahead = 1;  // gp_set_option("host_read_ahead", 0|1): read-ahead build of the one-launch host path
static int g_host_zero_copy = 1;   // gp_set_option("host_zero_copy", 0|1): one-launch host path of the specialised kernel
// gp_set_option("host_warps", n): warps per SM of the one-launch host path's kernel build.  The host path is
// bound by PCIe, not by the solve: a wave of warps uploads before anything downloads, so MORE warps per SM
// lengthen the part of a call in which only one PCIe direction is busy.  Measured on C2 (65,536 scenarios,
// V returned / |V| + losses returned, solves/s): 7 warps 5.85e7 / -, 6 warps 6.01e7 / 8.98e7,
// 4 warps 6.26e7 / 9.58e7, 3 warps 6.36e7 / 8.53e7.  Default 4; the device path uses every warp that fits.
static int g_host_warps = 4;
static int g_host_chunks = 0;  // gp_set_option("host_chunks", n): 0 = automatic
// gp_set_option("fbs_kernel", v): 0 automatic (feeder-specialised on-chip kernel once built, else two-sweep), 1 fused
// depth-first, 2 phase-parallel, 3 two-sweep streaming, 4 on-chip (error if the feeder does not fit),
// 5 feeder-specialised on-chip kernel (error unless gp_fbs_specialise succeeded)
static int g_fbs_kernel = 0;
// gp_set_option("fbs_prefetch", d): prefetch distance of the streaming kernel in steps; -1 (default) = automatic = 0.
// d in 1..64: last iteration's edge currents / the loads of event e + d are requested into L2; 100 + d: into L1, and
// the I / spu loads then look in L1 first.  Measured on B200, round 2 (depth-first order; profiles/r02f_*, r02m_*,
// r02n_*): with the final kernel build every distance costs time on both feeders - 123-bus 3.44 ms without, 3.63 ms
// into L1 at 1..3; 2,541-bus 61.4 ms without, 65.2-65.6 ms into L2 at 1, 2, 4, 64.1-64.2 ms into L1 - so it is off.
// (An earlier build of the same kernel gained 5 % on the 2,541-bus feeder from distance 2; the gain went away when
// the load path changed, which says how close to noise the effect is.)
static int g_fbs_prefetch = -1;
// gp_set_option("fbs_share_rows", 0|1): the streaming kernel skips the HBM round trips of I rows that only
// repeat a child's current or stay zero (default 1; 0 = every row travels, for A/B measurements)
static int g_fbs_share_rows = 1;
// gp_set_option("fbs_order", 0|1): 1 (default) = the streaming kernel walks the tree depth-first (forward step on
// the way down, backward step on the way up; same steps, same rows, same values), 0 = two full sweeps
static int g_fbs_order = 1;
// gp_set_option("fbs_walk_blocks", n): resident 128-thread blocks per SM the tree-walk kernel is compiled for
// (4 = up to 128 registers per thread, 5 = 96, 6 = 80, 8 = 64)
static int g_fbs_walk_blocks = 4;
static int g_fbs_auto_walk = 0;    // gp_set_option("fbs_auto_walk", 0|1): the automatic choice takes the tree-walk kernel

// |v| as numpy computes it for the ZIP model (abs then square, solution.py:202)
__device__ __forceinline__ double cabs2(double2 v, double& a) {
    a = sqrt(fma(v.x, v.x, v.y * v.y));
    return a * a;
}

// VoltVARController.get_Q (volt_var_controller.py:20-38) with the same operations in the same order: maxQ below V1,
// a ramp to 0 at V2, 0 up to V3, a ramp to minQ at V4, minQ above; the sign is flipped as the reference does.  The
// ramps' slope and intercept come from the host (same double arithmetic); slope * V + b is two roundings, not an FMA.
__device__ __forceinline__ double fbs_droop(double a, const FbsConst& k) {
    double q;
    if (a <= k.vv_v[0]) q = k.vv_q[1];
    else if (a <= k.vv_v[1]) q = __dadd_rn(__dmul_rn(k.vv_lo[0], a), k.vv_lo[1]);
    else if (a <= k.vv_v[2]) q = 0.0;
    else if (a <= k.vv_v[3]) q = __dadd_rn(__dmul_rn(k.vv_hi[0], a), k.vv_hi[1]);
    else q = k.vv_q[0];
    return -q;
}

// ---- where the state rows live --------------------------------------------------------------
// The sweep bodies below are written once against a row accessor:
//   GMem: rows in HBM, [row][ld] complex128 with the scenario innermost.  Row r of scenario s is
//         at base + s*16 + r*ldb (one 32x32->64 multiply-add per access).  State rows are read
//         through L2 only (ld.global.cg): they are touched once per sweep, so keeping them out of
//         L1 leaves L1 to the step table that every warp re-reads.
//   SMem: V and I rows in shared memory, [row][32 lanes] per warp (512-byte row pitch: lane l reads
//         bytes 16 l .. 16 l + 15 of a row, conflict-free 128-bit accesses); spu rows stay in HBM.
//   TAPS: the accessor also carries this scenario's column of the per-scenario regulator ratios
//         (GpSolveExtras.gamma_dev, [row][ld] doubles); the plain instance has no such member.
template <bool TAPS>
struct GMemT {
    static constexpr bool kTaps = TAPS;
    char* V;
    char* I;
    const char* S;
    unsigned ldb;
    const char* G;
    const char* W;      // this scenario's column of the per-scenario Solution.wpu rows ([V row][ld] doubles)
    const int* VV;      // [V row] non-zero = a volt-var controller sets wpu = get_Q(|V|) there at every calc_sV
    bool l1;            // prefetches go into L1 and the I / spu loads look there first (fbs_prefetch >= 100)
    __device__ __forceinline__ bool hasW() const { return TAPS && (W != nullptr || VV != nullptr); }
    // Solution.wpu of V row r whose voltage magnitude is a: the caller's row, or the droop of the row's controller
    __device__ __forceinline__ double wq(int r, double a, const FbsConst& k) const {
        double w = W ? __ldg(reinterpret_cast<const double*>(W + (size_t)(unsigned)r * (size_t)(ldb >> 1))) : 0.0;
        if (VV && __ldg(VV + r)) w = fbs_droop(a, k);
        return w;
    }
    __device__ __forceinline__ bool hasG() const { return TAPS && G != nullptr; }
    __device__ __forceinline__ double ldG(int r) const { return __ldg(reinterpret_cast<const double*>(G + (size_t)(unsigned)r * (size_t)(ldb >> 1))); }
    __device__ __forceinline__ double2 ldV(int r) const { return __ldcg(reinterpret_cast<const double2*>(V + (size_t)(unsigned)r * (size_t)ldb)); }
    __device__ __forceinline__ double2 ldI(int r) const {
        const double2* p = reinterpret_cast<const double2*>(I + (size_t)(unsigned)r * (size_t)ldb);
        return l1 ? __ldca(p) : __ldcg(p);
    }
    __device__ __forceinline__ double2 ldS(int r) const {
        const double2* p = reinterpret_cast<const double2*>(S + (size_t)(unsigned)r * (size_t)ldb);
        return l1 ? __ldca(p) : __ldcg(p);
    }
    __device__ __forceinline__ void stV(int r, double2 v) const { *reinterpret_cast<double2*>(V + (size_t)(unsigned)r * (size_t)ldb) = v; }
    __device__ __forceinline__ void stI(int r, double2 v) const { *reinterpret_cast<double2*>(I + (size_t)(unsigned)r * (size_t)ldb) = v; }
    // L2 prefetch of a row this thread will read a couple of steps from now (no register held)
    __device__ __forceinline__ void pf(const char* p) const {
        if (l1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        else asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
    __device__ __forceinline__ void pfV(int r) const { if (r >= 0) pf(V + (size_t)(unsigned)r * (size_t)ldb); }
    __device__ __forceinline__ void pfI(int r) const { if (r >= 0) pf(I + (size_t)(unsigned)r * (size_t)ldb); }
    __device__ __forceinline__ void pfS(int r) const { if (r >= 0) pf(S + (size_t)(unsigned)r * (size_t)ldb); }
};
using GMem = GMemT<false>;

struct SMem {
    static constexpr bool kTaps = false;
    __device__ __forceinline__ bool hasW() const { return false; }
    __device__ __forceinline__ double wq(int, double, const FbsConst&) const { return 0.0; }
    __device__ __forceinline__ bool hasG() const { return false; }
    __device__ __forceinline__ double ldG(int) const { return 0.0; }
    double2* V;         // this lane's column of the warp's V rows (shared memory)
    double2* I;
    const char* S;      // this scenario's column of the spu rows (global memory)
    unsigned ldb;
    __device__ __forceinline__ double2 ldV(int r) const { return V[r * 32]; }
    __device__ __forceinline__ double2 ldI(int r) const { return I[r * 32]; }
    __device__ __forceinline__ double2 ldS(int r) const { return __ldg(reinterpret_cast<const double2*>(S + (size_t)(unsigned)r * (size_t)ldb)); }
    __device__ __forceinline__ void stV(int r, double2 v) const { V[r * 32] = v; }
    __device__ __forceinline__ void stI(int r, double2 v) const { I[r * 32] = v; }
};

// ---- mask-specialised step bodies -----------------------------------------------------------
// Most buses carry exactly the phases of their incoming line and their parent carries them too.
// For those steps (FbsStep fast mask M != 0) the phase pattern is a compile-time constant: no
// row tests, no predicated arithmetic, straight-line FP64.  The branch on M is uniform (every
// thread walks the same feeder).  Everything else takes the general body.
// FIRST = the forward sweep of iteration 0: the currents are the flat start's zeros and are not read
// (V - Z 0 is computed all the same, so the values are those of the generic path bit for bit).
template <int M, class Mem, bool FIRST>
__device__ __forceinline__ void fwd_fast(const FbsStep* __restrict__ st, const int4 bv, const int4 pv, const Mem& m) {
    const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
    const int brow[3] = {bv.x, bv.y, bv.z}, prow[3] = {pv.x, pv.y, pv.z}, erow[3] = {ev.x, ev.y, ev.z};
    double2 Vp[3], Il[3];
#pragma unroll
    for (int p = 0; p < 3; ++p)
        if ((M >> p) & 1) {
            Vp[p] = m.ldV(prow[p]);
            Il[p] = FIRST ? make_double2(0.0, 0.0) : m.ldI(erow[p]);
        }
#pragma unroll
    for (int p = 0; p < 3; ++p)
        if ((M >> p) & 1) {
            double2 v = Vp[p];
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if ((M >> q) & 1) {
                    const double2 z = __ldg(&st->z[3 * p + q]);
                    v.x -= z.x * Il[q].x - z.y * Il[q].y;
                    v.y -= z.x * Il[q].y + z.y * Il[q].x;
                }
            m.stV(brow[p], v);
        }
}

template <int M, class Mem, bool UNUSED>
__device__ __forceinline__ void bwd_fast(const FbsStep* __restrict__ st, const int4 bv, const int4 pv,
                                         const int4* __restrict__ child_irow, const FbsConst& k, const Mem& m,
                                         const double2 (&sp)[3]) {
    const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
    const int brow[3] = {bv.x, bv.y, bv.z}, prow[3] = {pv.x, pv.y, pv.z};
    const int erow[3] = {ev.x, ev.y, ev.z};
    const int noload = (bv.w >> 22) & 7;        // bus phases without load and capacitor (fbs_tables_from_desc)
    double2 Vb[3], In[3];
#pragma unroll
    for (int p = 0; p < 3; ++p)
        if ((M >> p) & 1) {
            Vb[p] = m.ldV(brow[p]);
            In[p] = make_double2(0.0, 0.0);
        }
    for (int c = 0; c < ev.w; ++c) {                    // children's currents (:274-284)
        const int4 cr = __ldg(child_irow + pv.w + c);
        const int crow[3] = {cr.x, cr.y, cr.z};
#pragma unroll
        for (int p = 0; p < 3; ++p)
            if (((M >> p) & 1) && crow[p] >= 0) {
                const double2 ic = m.ldI(crow[p]);
                In[p].x += ic.x;
                In[p].y += ic.y;
            }
    }
    double sum = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
        if ((M >> p) & 1) {                             // calc_sV (solution.py:177-220) + update_current (:256-288)
            const double2 v = Vb[p];
            if (((noload >> p) & 1) && !m.hasW()) {
                // no load, no capacitor: conj(0 / V) adds an exact zero - unless V is not finite, when the
                // reference's 0 * inf makes the current NaN (nan_to_num then zeroes it): x - x does both
                const double t = (v.x - v.x) + (v.y - v.y);
                In[p].x += t;
                In[p].y += t;
            } else {
                double a;
                const double a2 = cabs2(v, a);
                const double f = k.aPQ + k.aI * a + k.aZ * a2;
                double2 sv = make_double2(sp[p].x * f, sp[p].y * f);
                sv.y -= __ldg(&st->cap[p]) * a2;
                if (m.hasW()) sv.y += m.wq(brow[p], a, k);      // + 1j * wpu (solution.py:203)
                if (v.x != 0.0 || v.y != 0.0) {
                    const double inv = 1.0 / fma(v.x, v.x, v.y * v.y);
                    In[p].x += (sv.x * v.x + sv.y * v.y) * inv;
                    In[p].y -= (sv.y * v.x - sv.x * v.y) * inv;
                }
            }
            sum += fabs(In[p].x) + fabs(In[p].y);
        }
    if (!(sum <= 1.7976931348623157e308)) {             // nan_to_num (:288) acts on non-finite values only
#pragma unroll
        for (int p = 0; p < 3; ++p)
            if ((M >> p) & 1) In[p] = make_double2(nan_to_num(In[p].x), nan_to_num(In[p].y));
    }
    const int no_store = (bv.w >> 16) & 7, no_pstore = (bv.w >> 19) & 7;
#pragma unroll
    for (int p = 0; p < 3; ++p)
        if (((M >> p) & 1) && !((no_store >> p) & 1)) m.stI(erow[p], In[p]);
#pragma unroll
    for (int p = 0; p < 3; ++p)
        if (((M >> p) & 1) && !((no_pstore >> p) & 1)) {    // update_voltage_backward (:207-232), surviving stores only
            double2 v = Vb[p];
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if ((M >> q) & 1) {
                    const double2 z = __ldg(&st->z[3 * p + q]);
                    v.x += z.x * In[q].x - z.y * In[q].y;
                    v.y += z.x * In[q].y + z.y * In[q].x;
                }
            m.stV(prow[p], v);
        }
}

#define GP_FAST_SWITCH(fn, X, ...)                 \
    switch (fast) {                                \
        case 7: fn<7, Mem, X>(__VA_ARGS__); return; \
        case 1: fn<1, Mem, X>(__VA_ARGS__); return; \
        case 2: fn<2, Mem, X>(__VA_ARGS__); return; \
        case 4: fn<4, Mem, X>(__VA_ARGS__); return; \
        case 3: fn<3, Mem, X>(__VA_ARGS__); return; \
        case 5: fn<5, Mem, X>(__VA_ARGS__); return; \
        case 6: fn<6, Mem, X>(__VA_ARGS__); return; \
        default: break;                            \
    }

// forward step of one bus: update_voltage_forward :162-184 / vr_forward :138-160
template <class Mem, bool FIRST = false>
__device__ __forceinline__ void fbs_fwd_step(const FbsStep* __restrict__ st, const Mem& m) {
    const double2 zero = make_double2(0.0, 0.0);
    const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_vrow));
    const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_vrow));
    const int kind = bv.w & 0xff, fast = (bv.w >> 8) & 0xff;
    GP_FAST_SWITCH(fwd_fast, FIRST, st, bv, pv, m)
    const int brow[3] = {bv.x, bv.y, bv.z};
    const int prow[3] = {pv.x, pv.y, pv.z};
    double2 Vp[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) Vp[p] = prow[p] >= 0 ? m.ldV(prow[p]) : zero;
    if (kind == GP_EDGE_VR) {   // vr_forward :138-160
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            double g = __ldg(&st->gamma[p]);
            if (g != 0.0 && brow[p] >= 0) {
                if (m.hasG()) g = m.ldG(__ldg(&st->gam_row[p]));        // this scenario's tap position
                m.stV(brow[p], make_double2(g * Vp[p].x, g * Vp[p].y));
            }
        }
        return;
    }
    const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
    const int erow[3] = {ev.x, ev.y, ev.z};
    double2 Il[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) Il[p] = (!FIRST && erow[p] >= 0) ? m.ldI(erow[p]) : zero;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (brow[p] < 0) continue;
        double2 v = Vp[p];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (erow[q] < 0) continue;
            const double2 z = __ldg(&st->z[3 * p + q]);
            v.x -= z.x * Il[q].x - z.y * Il[q].y;
            v.y -= z.x * Il[q].y + z.y * Il[q].x;
        }
        m.stV(brow[p], v);
    }
}

// spu rows of the bus of a step (zeros where the bus-phase carries no load)
template <class Mem>
__device__ __forceinline__ void fbs_load_spu(const FbsStep* __restrict__ st, const Mem& m, double2 (&sp)[3]) {
    const int4 lv = __ldg(reinterpret_cast<const int4*>(st->load_srow));
    const int lrow[3] = {lv.x, lv.y, lv.z};
#pragma unroll
    for (int p = 0; p < 3; ++p) sp[p] = lrow[p] >= 0 ? m.ldS(lrow[p]) : make_double2(0.0, 0.0);
}

// backward step of one bus (:104-121): calc_sV, update_current, update_voltage_backward / vr_backward
template <class Mem>
__device__ __forceinline__ void fbs_bwd_step(const FbsStep* __restrict__ st, const int4* __restrict__ child_irow,
                                             const FbsConst& k, const Mem& m, const double2 (&sp)[3]) {
    const double2 zero = make_double2(0.0, 0.0);
    const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_vrow));
    const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_vrow));
    const int kind = bv.w & 0xff, fast = (bv.w >> 8) & 0xff;
    const int no_store = (bv.w >> 16) & 7;      // phases whose I row is a shared row: value in place already
    // phases whose parent-voltage rewrite a sibling earlier in topo order overwrites before anything reads it
    // (fbs_mark_dead_parent_stores): only the store that survives is made
    const int no_pstore = (bv.w >> 19) & 7;
    GP_FAST_SWITCH(bwd_fast, false, st, bv, pv, child_irow, k, m, sp)
    const int brow[3] = {bv.x, bv.y, bv.z};
    const int prow[3] = {pv.x, pv.y, pv.z};
    const int child_begin = pv.w;
    double2 Vb[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) Vb[p] = brow[p] >= 0 ? m.ldV(brow[p]) : zero;
    if (kind == GP_EDGE_VR) {   // vr_current leaves Itx = 0; vr_backward :186-205
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const double g = __ldg(&st->gamma[p]);
            if (g != 0.0 && prow[p] >= 0 && !((no_pstore >> p) & 1)) {
                double ig = __ldg(&st->inv_gamma[p]);
                if (m.hasG()) ig = 1.0 / m.ldG(__ldg(&st->gam_row[p]));     // 1/gamma * V (:205)
                m.stV(prow[p], make_double2(ig * Vb[p].x, ig * Vb[p].y));
            }
        }
        return;
    }
    const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
    const int erow[3] = {ev.x, ev.y, ev.z};
    const int child_count = ev.w;
    // calc_sV(bus) (solution.py:177-220) + update_current (:256-288)
    double2 In[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        In[p] = zero;
        if (brow[p] < 0) continue;
        const double2 v = Vb[p];
        if (((bv.w >> (22 + p)) & 1) && !m.hasW()) {    // no load, no capacitor: an exact zero, NaN if V is not finite
            const double t = (v.x - v.x) + (v.y - v.y);
            In[p] = make_double2(t, t);
            continue;
        }
        double a;
        const double a2 = cabs2(v, a);
        const double f = k.aPQ + k.aI * a + k.aZ * a2;
        double2 sv = make_double2(sp[p].x * f, sp[p].y * f);   // sp = 0 where there is no load
        sv.y -= __ldg(&st->cap[p]) * a2;
        if (m.hasW()) sv.y += m.wq(brow[p], a, k);              // + 1j * wpu (solution.py:203)
        if (v.x != 0.0 || v.y != 0.0) {
            // conj(sV / V) computed as (sV * conj(V)) / |V|^2, conjugated
            const double inv = 1.0 / fma(v.x, v.x, v.y * v.y);
            const double qr = (sv.x * v.x + sv.y * v.y) * inv;
            const double qi = (sv.y * v.x - sv.x * v.y) * inv;
            In[p] = make_double2(qr, -qi);
        }
    }
    for (int c = 0; c < child_count; ++c) {
        const int4 cr4 = __ldg(child_irow + child_begin + c);
        const int cr[3] = {cr4.x, cr4.y, cr4.z};
#pragma unroll
        for (int p = 0; p < 3; ++p)
            if (cr[p] >= 0) {
                const double2 ic = m.ldI(cr[p]);
                In[p].x += ic.x;
                In[p].y += ic.y;
            }
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (erow[p] >= 0) {
            In[p] = make_double2(nan_to_num(In[p].x), nan_to_num(In[p].y));
            if (!((no_store >> p) & 1)) m.stI(erow[p], In[p]);
        } else {
            In[p] = zero;
        }
    }
    // update_voltage_backward (:207-232): phases of the child bus, masked by the parent
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        if (brow[p] < 0 || prow[p] < 0 || ((no_pstore >> p) & 1)) continue;
        double2 v = Vb[p];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (erow[q] < 0) continue;
            const double2 z = __ldg(&st->z[3 * p + q]);
            v.x += z.x * In[q].x - z.y * In[q].y;
            v.y += z.x * In[q].y + z.y * In[q].x;
        }
        m.stV(prow[p], v);
    }
}

// root mismatch (:123-128): max_ph |V[root] - Vref|, NaN-propagating like numpy's max
template <class Mem>
__device__ __forceinline__ double fbs_root_diff(const FbsConst& k, const Mem& m) {
    double diff = 0.0;
    bool nan = false;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const double2 v = m.ldV(k.root_vrow[p]);
        const double dx = v.x - k.vref[p].x, dy = v.y - k.vref[p].y;
        const double d = sqrt(fma(dx, dx, dy * dy));
        nan |= (d != d);
        diff = fmax(diff, d);
    }
    return nan ? __longlong_as_double(0x7ff8000000000000LL) : diff;
}

__device__ __forceinline__ int fbs_status(bool converged, double diff) {
    return converged ? GP_ST_CONVERGED : ((diff != diff || isinf(diff)) ? GP_ST_NONFINITE : GP_ST_MAXITER);
}

// EX = the instance behind gp_fbs_solve_batch_ex: per-scenario regulator ratios (gam != NULL) and / or a
// warm start (V and I hold an earlier solution).  The plain instance carries neither argument.
// DFS = the steps run in depth-first order (fbs_order 1, the default), else as two full sweeps.
template <int BLOCK, bool EX, bool DFS>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK)   // <= 64 registers: 1024 resident threads per SM
fbs_solve_kernel(const FbsStep* __restrict__ steps, const int4* __restrict__ child_irow,
                 FbsConst k, int n_steps, int n_vrows, int n_irows, long long S, unsigned ldb,
                 const char* __restrict__ spu, char* __restrict__ V,
                 char* __restrict__ I, int* __restrict__ iters_out,
                 int* __restrict__ status_out, double* __restrict__ diff_out, double tol,
                 int maxiter, int PF, const int* __restrict__ unfed, int n_unfed,
                 const char* __restrict__ gam, int warm, const char* __restrict__ wpu,
                 const int2* __restrict__ alias, int n_alias, const int2* __restrict__ events, int n_events,
                 const int* __restrict__ vv) {
    const long long s = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (s >= S) return;
    typedef GMemT<EX> GMem;
    const GMem m{V + s * 16, I + s * 16, spu + s * 16, ldb, (EX && gam) ? gam + s * 8 : nullptr,
                 (EX && wpu) ? wpu + s * 8 : nullptr, EX ? vv : nullptr, PF >= 100};
    if (PF >= 100) PF -= 100;
    const double2 zero = make_double2(0.0, 0.0);
    if (EX && warm) n_unfed = 0;    // warm start: every V and I row keeps the earlier solution

    // flat start (solution.py:108-125): V = I = 0.  Every I row is written by the first backward
    // sweep before anything reads it again, and the first forward sweep knows I = 0 without
    // reading it; only V rows that no forward step writes (non-regulated phases below a
    // regulator) must hold their initial zero.  Zeroing just those saves one write of every
    // V and I row and one read of every I row per solve.
    for (int j = 0; j < n_unfed; ++j) m.stV(__ldg(unfed + j), zero);
    if (maxiter < 1 && !(EX && warm)) {   // no sweep at all: the outputs are the flat start itself
        for (int r = 0; r < n_vrows; ++r) m.stV(r, zero);
        for (int r = 0; r < n_irows; ++r) m.stI(r, zero);
    }
    int it = 0;
    double diff = -1.0;
    bool converged = false;   // Vtest = 0 -> max|0 - Vref| = 1 > tol (solution_fbs.py:85)
    // The rows a step reads that were last written a whole sweep ago (the I rows of its edge in
    // the forward sweep, the V and spu rows of its bus in the backward sweep) come from DRAM.
    // Their addresses depend only on the step table, so they are prefetched into L2 PF steps
    // ahead: the dependent part of a step then waits for an L2 hit instead of a DRAM access.

    while (!converged && it < maxiter) {
#pragma unroll
        for (int p = 0; p < 3; ++p) m.stV(k.root_vrow[p], k.vref[p]);      // V[root] = Vref  (:89)
        if (DFS) {
            // Depth-first order of the SAME steps on the same rows (gp_set_option("fbs_order")).  A bus's forward
            // step runs when the walk descends its edge, its backward step when the walk comes back up, children
            // latest-in-topo-order first.  Every value a step reads is the one the reference's two full sweeps
            // give it: a forward step needs its parent's FORWARD voltage, and of a bus's children only the one
            // earliest in topo order - visited last here - rewrites a phase of the parent (the other rewrites
            // are overwritten before anything reads them and are not made: no_pstore); a backward step needs
            // its children's new currents and its own bus as they rewrote it, all done when the walk returns;
            // the edge current a forward step reads is last iteration's, its row is only written on the way up.
            // What changes is WHEN rows are touched: a bus's voltage is written and read again within its own
            // subtree, for most buses a few steps apart, so the backward step finds it in L2 instead of DRAM.
            const bool first = it == 0 && !(EX && warm);
            for (int e = 0; e < n_events; ++e) {
                const int2 ev = __ldg(events + e);
                if (PF > 0 && e + PF < n_events) {
                    // the rows of event e + PF that were last written a whole iteration ago (the edge currents of a
                    // forward step) or never by this kernel (the loads of a backward step) come from DRAM: into L2 now
                    const int2 pe = __ldg(events + e + PF);
                    const FbsStep* ps = steps + pe.x;
                    if (pe.y & 1) {
                        const int4 lv = __ldg(reinterpret_cast<const int4*>(ps->load_srow));
                        m.pfS(lv.x);
                        m.pfS(lv.y);
                        m.pfS(lv.z);
                    } else if (!first) {
                        const int4 ei = __ldg(reinterpret_cast<const int4*>(ps->edge_irow));
                        m.pfI(ei.x);
                        m.pfI(ei.y);
                        m.pfI(ei.z);
                    }
                }
                if (ev.y & 1) {
                    double2 sp[3];
                    fbs_load_spu(steps + ev.x, m, sp);
                    fbs_bwd_step(steps + ev.x, child_irow, k, m, sp);
                } else if (first) {
                    fbs_fwd_step<GMem, true>(steps + ev.x, m);
                } else {
                    fbs_fwd_step<GMem, false>(steps + ev.x, m);
                }
            }
        } else {
        for (int pos = 0; pos < n_steps; ++pos) {                                      // :92-98
            if (PF > 0 && pos + PF < n_steps) {
                const int4 ev = __ldg(reinterpret_cast<const int4*>(steps[pos + PF].edge_irow));
                m.pfI(ev.x);
                m.pfI(ev.y);
                m.pfI(ev.z);
            }
            if (it == 0 && !(EX && warm)) fbs_fwd_step<GMem, true>(steps + pos, m);
            else fbs_fwd_step<GMem, false>(steps + pos, m);
        }
        for (int pos = n_steps - 1; pos >= 0; --pos) {                                 // :104-121
            if (PF > 0 && pos >= PF) {
                const int4 bv = __ldg(reinterpret_cast<const int4*>(steps[pos - PF].bus_vrow));
                const int4 lv = __ldg(reinterpret_cast<const int4*>(steps[pos - PF].load_srow));
                m.pfV(bv.x);
                m.pfV(bv.y);
                m.pfV(bv.z);
                m.pfS(lv.x);
                m.pfS(lv.y);
                m.pfS(lv.z);
            }
            double2 sp[3];
            fbs_load_spu(steps + pos, m, sp);
            fbs_bwd_step(steps + pos, child_irow, k, m, sp);
        }
        }
        ++it;
        diff = fbs_root_diff(k, m);
        converged = diff <= tol;
    }
    // shared I rows (steps = the d_steps_a tables): the caller's array gets every row, once
    for (int j = 0; j < n_alias; ++j) {
        const int2 a = __ldg(alias + j);
        m.stI(a.x, a.y >= 0 ? m.ldI(a.y) : zero);
    }
    iters_out[s] = it;
    diff_out[s] = diff;
    status_out[s] = fbs_status(converged, diff);
}

// ---------------------------------------------------------------------------------------
// On-chip two-sweep kernel: the whole solve of a scenario runs out of shared memory.
//
// For a small feeder the V and I rows of 32 scenarios (one warp) fit in a slice of the SM's
// 227 KB of shared memory, [row][lane] with a 512-byte pitch.  A warp owns its slice outright -
// a lane only ever touches its own column - so there is no block-level synchronisation at all.
// Per solve the HBM traffic drops from it * 16 (2 P_l + 3 P_b + P_L) to the spu reads
// (it * 16 P_L, L2-resident after the first sweep) plus ONE write of the final V and I rows:
// the sweep-to-sweep dependency chain runs at shared-memory latency instead of an L2 round trip
// per tree step, and the kernel becomes FP64-pipe / issue bound instead of HBM-latency bound.
//
// Scheduling: one persistent block per SM, NW warps.  Every warp pulls 32-scenario tiles from a
// global counter until the batch is exhausted (iteration counts differ between tiles, so static
// assignment would leave SMs idle); the last warp to finish re-arms the counter for the next
// launch.  The spu rows a backward step needs are fetched two steps ahead into registers.
// ---------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
fbs_onchip_kernel(const FbsStep* __restrict__ steps, const int4* __restrict__ child_irow, FbsConst k, int n_steps,
                  int n_vrows, int n_irows, long long S, unsigned ldb, const char* __restrict__ spu,
                  char* __restrict__ V, char* __restrict__ I, int* __restrict__ iters_out,
                  int* __restrict__ status_out, double* __restrict__ diff_out, double tol, int maxiter,
                  unsigned* __restrict__ sched) {
    extern __shared__ __align__(16) double2 gp_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* Vs = gp_smem + (size_t)warp * (size_t)(n_vrows + n_irows) * 32 + lane;
    double2* Is = Vs + (size_t)n_vrows * 32;
    const long long n_tiles = (S + 31) >> 5;
    const double2 zero = make_double2(0.0, 0.0);
    for (;;) {
        unsigned tile = 0;
        if (lane == 0) tile = atomicAdd(&sched[0], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if ((long long)tile >= n_tiles) break;
        const long long s_raw = (long long)tile * 32 + lane;
        const bool live = s_raw < S;
        const long long s = live ? s_raw : S - 1;          // idle lanes of the last tile shadow a real scenario
        const SMem m{Vs, Is, spu + s * 16, ldb};
        for (int r = 0; r < n_vrows; ++r) m.stV(r, zero);  // flat start (solution.py:108-125)
        for (int r = 0; r < n_irows; ++r) m.stI(r, zero);
        int it = 0;
        double diff = -1.0;
        bool converged = false;
        while (!converged && it < maxiter) {
#pragma unroll
            for (int p = 0; p < 3; ++p) m.stV(k.root_vrow[p], k.vref[p]);
            // spu rows of the first two backward steps: in flight during the forward sweep
            double2 spA[3], spB[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) spA[p] = spB[p] = zero;
            if (n_steps > 0) fbs_load_spu(steps + (n_steps - 1), m, spA);
            if (n_steps > 1) fbs_load_spu(steps + (n_steps - 2), m, spB);
            for (int pos = 0; pos < n_steps; ++pos) fbs_fwd_step(steps + pos, m);
            for (int pos = n_steps - 1; pos >= 0; --pos) {
                double2 spC[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) spC[p] = zero;
                if (pos >= 2) fbs_load_spu(steps + (pos - 2), m, spC);
                fbs_bwd_step(steps + pos, child_irow, k, m, spA);
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    spA[p] = spB[p];
                    spB[p] = spC[p];
                }
            }
            ++it;
            diff = fbs_root_diff(k, m);
            converged = diff <= tol;
        }
        __syncwarp();
        if (live) {         // final V and I rows: one coalesced 512-byte store per row and warp
            char* Vg = V + s * 16;
            char* Ig = I + s * 16;
            for (int r = 0; r < n_vrows; ++r)
                *reinterpret_cast<double2*>(Vg + (size_t)(unsigned)r * (size_t)ldb) = m.ldV(r);
            for (int r = 0; r < n_irows; ++r)
                *reinterpret_cast<double2*>(Ig + (size_t)(unsigned)r * (size_t)ldb) = m.ldI(r);
            iters_out[s] = it;
            diff_out[s] = diff;
            status_out[s] = fbs_status(converged, diff);
        }
    }
    if (lane == 0) {        // the last warp out re-arms the tile counter for the next launch
        __threadfence();
        const unsigned done = atomicAdd(&sched[1], 1u);
        if (done == gridDim.x * NW - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------
// Phase-parallel two-sweep kernel.
//
// Same sweeps as fbs_solve_kernel, but a scenario is owned by THREE lanes, one per phase
// (lane = phase * 10 + j inside a warp: 10 scenarios per warp, lanes 30/31 idle).  A lane
// touches only the rows of its own phase, so there is no cross-lane memory dependency; the
// 3x3 coupling (Z I) is the only exchange and goes through warp shuffles.  Per warp a row
// access is three 160-byte coalesced segments.  Compared with one thread per scenario the
// serial chain per tree step is a third as long and there are three times as many warps to
// hide latency with, which is what a batch of about one wave (BASELINE config C2) needs.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int sel3(const int4 v, int g) { return g == 0 ? v.x : (g == 1 ? v.y : v.z); }

__device__ __forceinline__ double2 shfl2(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1536 / BLOCK)   // <= 42 registers: 48 warps per SM
fbs_phase_kernel(const FbsStep* __restrict__ steps, const int4* __restrict__ child_irow, FbsConst k,
                 int n_steps, int n_vrows, int n_irows, long long S, unsigned ldb,
                 const char* __restrict__ spu, char* __restrict__ V, char* __restrict__ I,
                 int* __restrict__ iters_out, int* __restrict__ status_out, double* __restrict__ diff_out,
                 double tol, int maxiter) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * BLOCK + threadIdx.x) >> 5;
    const int j = lane % 10;
    const int g = lane < 30 ? lane / 10 : 2;            // phase of this lane (idle lanes mirror phase c)
    const long long s_raw = warp * 10 + j;
    const bool owner = lane < 30 && s_raw < S;          // lanes that own real data
    const long long s = s_raw < S ? s_raw : S - 1;
    const int src1 = ((g + 1) % 3) * 10 + j, src2 = ((g + 2) % 3) * 10 + j;   // the other two phases
    const int g1 = (g + 1) % 3, g2 = (g + 2) % 3;
    const char* spu_b = spu + s * 16;
    char* V_b = V + s * 16;
    char* I_b = I + s * 16;
#define VLOAD(r) __ldcg(reinterpret_cast<const double2*>(V_b + (size_t)(unsigned)(r) * (size_t)ldb))
#define ILOAD(r) __ldcg(reinterpret_cast<const double2*>(I_b + (size_t)(unsigned)(r) * (size_t)ldb))
#define SLOAD(r) __ldcg(reinterpret_cast<const double2*>(spu_b + (size_t)(unsigned)(r) * (size_t)ldb))
#define VSTORE(r, v) (*reinterpret_cast<double2*>(V_b + (size_t)(unsigned)(r) * (size_t)ldb) = (v))
#define ISTORE(r, v) (*reinterpret_cast<double2*>(I_b + (size_t)(unsigned)(r) * (size_t)ldb) = (v))
    const double2 zero = make_double2(0.0, 0.0);
    if (owner) {     // flat start: V = I = 0 (each phase lane clears a third of the rows)
        for (int r = g; r < n_vrows; r += 3) VSTORE(r, zero);
        for (int r = g; r < n_irows; r += 3) ISTORE(r, zero);
    }
    __syncwarp();
    const int root_row = g == 0 ? k.root_vrow[0] : (g == 1 ? k.root_vrow[1] : k.root_vrow[2]);
    const double2 vref = g == 0 ? k.vref[0] : (g == 1 ? k.vref[1] : k.vref[2]);
    int it = 0;
    double diff = -1.0;
    bool converged = false;
    bool active = owner;
    while (__any_sync(0xffffffffu, active)) {
        if (active) VSTORE(root_row, vref);                                  // V[root] = Vref (:89)
        __syncwarp();        // the zero-fill above was striped over the three lanes of a scenario
        // ---- forward sweep (:92-98) -----------------------------------------------------
        for (int pos = 0; pos < n_steps; ++pos) {
            const FbsStep* st = steps + pos;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_vrow));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_vrow));
            const int myb = sel3(bv, g), myp = sel3(pv, g);
            const double2 Vp = (active && myp >= 0) ? VLOAD(myp) : zero;
            if ((bv.w & 0xff) == GP_EDGE_VR) {                               // vr_forward :138-160
                const double gm = __ldg(&st->gamma[g]);
                if (active && gm != 0.0 && myb >= 0) VSTORE(myb, make_double2(gm * Vp.x, gm * Vp.y));
                continue;
            }
            const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
            const int mye = sel3(ev, g);
            const double2 I0 = (active && mye >= 0) ? ILOAD(mye) : zero;
            const double2 I1 = shfl2(I0, src1), I2 = shfl2(I0, src2);
            if (active && myb >= 0) {                                        // update_voltage_forward :162-184
                double2 v = Vp;
                const double2 z0 = __ldg(&st->z[3 * g + g]), z1 = __ldg(&st->z[3 * g + g1]),
                              z2 = __ldg(&st->z[3 * g + g2]);
                v.x -= z0.x * I0.x - z0.y * I0.y;
                v.y -= z0.x * I0.y + z0.y * I0.x;
                v.x -= z1.x * I1.x - z1.y * I1.y;
                v.y -= z1.x * I1.y + z1.y * I1.x;
                v.x -= z2.x * I2.x - z2.y * I2.y;
                v.y -= z2.x * I2.y + z2.y * I2.x;
                VSTORE(myb, v);
            }
        }
        // ---- backward sweep (:104-121) ---------------------------------------------------
        for (int pos = n_steps - 1; pos >= 0; --pos) {
            const FbsStep* st = steps + pos;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_vrow));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_vrow));
            const int myb = sel3(bv, g), myp = sel3(pv, g);
            const double2 Vb = (active && myb >= 0) ? VLOAD(myb) : zero;
            if ((bv.w & 0xff) == GP_EDGE_VR) {                               // vr_backward :186-205
                const double gm = __ldg(&st->gamma[g]);
                if (active && gm != 0.0 && myp >= 0) {
                    const double ig = __ldg(&st->inv_gamma[g]);
                    VSTORE(myp, make_double2(ig * Vb.x, ig * Vb.y));
                }
                continue;
            }
            const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
            const int4 lv = __ldg(reinterpret_cast<const int4*>(st->load_srow));
            const int mye = sel3(ev, g), myl = sel3(lv, g);
            const double2 sp = (active && myl >= 0) ? SLOAD(myl) : zero;
            double2 In = zero;
            for (int c = 0; c < ev.w; ++c) {                                 // children's currents (:274-284)
                const int r = sel3(__ldg(child_irow + pv.w + c), g);
                if (active && r >= 0) {
                    const double2 ic = ILOAD(r);
                    In.x += ic.x;
                    In.y += ic.y;
                }
            }
            if (myb >= 0) {                                                  // calc_sV + update_current
                double a;
                const double a2 = cabs2(Vb, a);
                const double f = k.aPQ + k.aI * a + k.aZ * a2;
                double2 sv = make_double2(sp.x * f, sp.y * f);
                sv.y -= __ldg(&st->cap[g]) * a2;
                if (Vb.x != 0.0 || Vb.y != 0.0) {
                    const double inv = 1.0 / fma(Vb.x, Vb.x, Vb.y * Vb.y);
                    In.x += (sv.x * Vb.x + sv.y * Vb.y) * inv;
                    In.y -= (sv.y * Vb.x - sv.x * Vb.y) * inv;
                }
            }
            if (mye >= 0) {
                In = make_double2(nan_to_num(In.x), nan_to_num(In.y));
                if (active) ISTORE(mye, In);
            } else {
                In = zero;
            }
            const double2 I1 = shfl2(In, src1), I2 = shfl2(In, src2);
            if (active && myb >= 0 && myp >= 0) {                            // update_voltage_backward :207-232
                double2 v = Vb;
                const double2 z0 = __ldg(&st->z[3 * g + g]), z1 = __ldg(&st->z[3 * g + g1]),
                              z2 = __ldg(&st->z[3 * g + g2]);
                v.x += z0.x * In.x - z0.y * In.y;
                v.y += z0.x * In.y + z0.y * In.x;
                v.x += z1.x * I1.x - z1.y * I1.y;
                v.y += z1.x * I1.y + z1.y * I1.x;
                v.x += z2.x * I2.x - z2.y * I2.y;
                v.y += z2.x * I2.y + z2.y * I2.x;
                VSTORE(myp, v);
            }
        }
        // ---- convergence (:123-128): max over the three phase lanes ---------------------------
        double d = 0.0;
        if (active) {
            const double2 v = VLOAD(root_row);
            const double dx = v.x - vref.x, dy = v.y - vref.y;
            d = sqrt(fma(dx, dx, dy * dy));
        }
        const double d1 = __shfl_sync(0xffffffffu, d, src1), d2 = __shfl_sync(0xffffffffu, d, src2);
        if (active) {
            ++it;
            const bool nan = (d != d) | (d1 != d1) | (d2 != d2);
            diff = nan ? __longlong_as_double(0x7ff8000000000000LL) : fmax(d, fmax(d1, d2));
            converged = diff <= tol;
            active = !converged && it < maxiter;
        }
    }
#undef VLOAD
#undef ILOAD
#undef SLOAD
#undef VSTORE
#undef ISTORE
    if (owner && lane < 10) {
        iters_out[s] = it;
        diff_out[s] = diff;
        status_out[s] = converged ? GP_ST_CONVERGED : ((diff != diff || isinf(diff)) ? GP_ST_NONFINITE : GP_ST_MAXITER);
    }
}


// ---------------------------------------------------------------------------------------
// Fused depth-first sweep.
//
// The reference runs a full forward sweep, then a full backward sweep.  Per scenario the
// two can be fused into ONE depth-first walk of the tree without changing any value:
// descending an edge is that edge's forward update (it needs only the parent's FORWARD
// voltage and last iteration's edge current), ascending it is the bus's backward step
// (load, current, parent-voltage rewrite), which needs the bus voltage after all its
// children rewrote it and the children's new currents.  Children are visited latest-in-
// topo-order first, so the per-phase overwrite of the parent voltage ends with the same
// "winning child" (earliest in topo_order) as the reference's reversed(topo_order) loop.
//
// The walk keeps the path's forward voltage, rewritten voltage and child-current sum in
// registers for the current bus and in a per-thread stack (local memory, L1-resident)
// for its ancestors.  HBM traffic per sweep drops from 16(2 P_l + 3 P_b + P_L) to
// 16(2 P_l + P_b + P_L): I_old read, I_new written, spu read, final V written; the
// voltage no longer round-trips between the sweeps and the dependent parent->child chain
// runs through registers/L1 instead of L2.
// ---------------------------------------------------------------------------------------
constexpr int EV_ASCEND = 1, EV_VF = 2, EV_VBIA = 4;   // event flags; level = flags >> 4
constexpr int DFS_PF = 4;                               // prefetch distance in events

struct PathState {
    double2 Vf[3];   // forward voltage of the bus
    double2 Vb[3];   // voltage as rewritten by the children processed so far
    double2 Ia[3];   // sum of the new currents of those children
};

template <int BLOCK, int MAXD>
__global__ void __launch_bounds__(BLOCK)
fbs_dfs_kernel(const FbsStep* __restrict__ steps, const int2* __restrict__ events,
               const int4* __restrict__ evrows, int n_events,
               FbsConst k, long long S, unsigned ldb, const char* __restrict__ spu, char* __restrict__ V,
               char* __restrict__ I, int* __restrict__ iters_out, int* __restrict__ status_out,
               double* __restrict__ diff_out, double tol, int maxiter) {
    const long long s = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (s >= S) return;
    const char* spu_s = spu + s * 16;
    char* Vs = V + s * 16;
    char* Is = I + s * 16;
#define GP_ROW(base, row) (reinterpret_cast<double2*>((base) + (size_t)(unsigned)(row) * (size_t)ldb))
#define GP_CROW(base, row) (reinterpret_cast<const double2*>((base) + (size_t)(unsigned)(row) * (size_t)ldb))
    PathState stk[MAXD];
    // Per-thread ring of asynchronously prefetched operand rows (cp.async / LDGSTS), DFS_PF
    // events ahead: the I_old rows of a descend, the spu rows of an ascend.  Neither depends on
    // what the walk computes meanwhile, so the HBM latency of event e + DFS_PF overlaps the
    // arithmetic of event e without holding registers.  Layout [slot][thread]: conflict-free.
    __shared__ double2 ring[DFS_PF * 3][BLOCK];
    const int tid = threadIdx.x;
    int it = 0;
    double diff = -1.0;
    bool converged = false;
    auto issue = [&](int e) {
        if (e < n_events) {
            const int4 r = __ldg(evrows + e);
            const int rows[3] = {r.x, r.y, r.z};
            const char* base = r.w ? Is : spu_s;
            if (!(r.w && it == 0)) {                 // flat start: I_old = 0, nothing to fetch
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    if (rows[p] >= 0) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring[(e % DFS_PF) * 3 + p][tid]);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst),
                                     "l"(base + (size_t)(unsigned)rows[p] * (size_t)ldb)
                                     : "memory");
                    }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    while (!converged && it < maxiter) {
        PathState cur;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            cur.Vf[p] = cur.Vb[p] = k.vref[p];                      // V[root] = Vref (:89)
            cur.Ia[p] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int e = 0; e < DFS_PF; ++e) issue(e);
        for (int e = 0; e < n_events; ++e) {
            const int2 ev = __ldg(events + e);
            // operands of this event have landed; take them and reuse the slot for e + DFS_PF
            asm volatile("cp.async.wait_group %0;" ::"n"(DFS_PF - 1) : "memory");
            double2 D[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) D[p] = ring[(e % DFS_PF) * 3 + p][tid];
            issue(e + DFS_PF);
            const FbsStep* st = steps + ev.x;
            const int L = ev.y >> 4;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_vrow));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_vrow));
            const int brow[3] = {bv.x, bv.y, bv.z};
            const int prow[3] = {pv.x, pv.y, pv.z};
            const int kind = bv.w & 0xff;
            if ((ev.y & EV_ASCEND) == 0) {
                // ---------------- descend: forward update of this edge (:162-184 / :138-160)
                // Save only what a later event of the parent still needs (see build_events).
                if (ev.y & EV_VF) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) stk[L - 1].Vf[p] = cur.Vf[p];
                }
                if (ev.y & EV_VBIA) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        stk[L - 1].Vb[p] = cur.Vb[p];
                        stk[L - 1].Ia[p] = cur.Ia[p];
                    }
                }
                double2 Vc[3];
                if (kind == GP_EDGE_VR) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const double g = __ldg(&st->gamma[p]);
                        Vc[p] = (brow[p] >= 0) ? make_double2(g * cur.Vf[p].x, g * cur.Vf[p].y) : make_double2(0.0, 0.0);
                    }
                } else {
                    const int4 er = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
                    const int erow[3] = {er.x, er.y, er.z};
                    double2 Il[3];
#pragma unroll
                    for (int p = 0; p < 3; ++p)
                        Il[p] = (it > 0 && erow[p] >= 0) ? D[p] : make_double2(0.0, 0.0);
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        double2 v = make_double2(0.0, 0.0);
                        if (brow[p] >= 0) {
                            v = cur.Vf[p];
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                if (erow[q] < 0) continue;
                                const double2 z = __ldg(&st->z[3 * p + q]);
                                v.x -= z.x * Il[q].x - z.y * Il[q].y;
                                v.y -= z.x * Il[q].y + z.y * Il[q].x;
                            }
                        }
                        Vc[p] = v;
                    }
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    cur.Vf[p] = cur.Vb[p] = Vc[p];
                    cur.Ia[p] = make_double2(0.0, 0.0);
                }
                continue;
            }
            // -------------------- ascend: backward step of this bus (:104-121)
            PathState par;
#pragma unroll
            for (int p = 0; p < 3; ++p) par.Vf[p] = par.Vb[p] = par.Ia[p] = make_double2(0.0, 0.0);
            if (ev.y & EV_VF) {
#pragma unroll
                for (int p = 0; p < 3; ++p) par.Vf[p] = stk[L - 1].Vf[p];
            }
            if (ev.y & EV_VBIA) {
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    par.Vb[p] = stk[L - 1].Vb[p];
                    par.Ia[p] = stk[L - 1].Ia[p];
                }
            } else {        // first child of its parent: nothing rewritten / summed yet
#pragma unroll
                for (int p = 0; p < 3; ++p) par.Vb[p] = par.Vf[p];
            }
#pragma unroll
            for (int p = 0; p < 3; ++p)
                if (brow[p] >= 0) *GP_ROW(Vs, brow[p]) = cur.Vb[p];
            if (kind == GP_EDGE_VR) {   // vr_current leaves Itx = 0; vr_backward :186-205
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const double g = __ldg(&st->gamma[p]);
                    if (g != 0.0 && prow[p] >= 0) {
                        const double ig = __ldg(&st->inv_gamma[p]);
                        par.Vb[p] = make_double2(ig * cur.Vb[p].x, ig * cur.Vb[p].y);
                    }
                }
                cur = par;
                continue;
            }
            const int4 er = __ldg(reinterpret_cast<const int4*>(st->edge_irow));
            const int4 lv = __ldg(reinterpret_cast<const int4*>(st->load_srow));
            const int erow[3] = {er.x, er.y, er.z};
            const int lrow[3] = {lv.x, lv.y, lv.z};
            double2 In[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                In[p] = make_double2(0.0, 0.0);
                if (erow[p] < 0) continue;
                double2 cu = cur.Ia[p];
                if (brow[p] >= 0) {      // calc_sV (solution.py:177-220) + update_current (:256-288)
                    const double2 v = cur.Vb[p];
                    double a;
                    const double a2 = cabs2(v, a);
                    double2 sv = make_double2(0.0, 0.0);
                    if (lrow[p] >= 0) {
                        const double2 sp = D[p];
                        const double f = k.aPQ + k.aI * a + k.aZ * a2;
                        sv.x = sp.x * f;
                        sv.y = sp.y * f;
                    }
                    sv.y -= __ldg(&st->cap[p]) * a2;
                    if (v.x != 0.0 || v.y != 0.0) {
                        const double inv = 1.0 / fma(v.x, v.x, v.y * v.y);
                        cu.x += (sv.x * v.x + sv.y * v.y) * inv;
                        cu.y -= (sv.y * v.x - sv.x * v.y) * inv;
                    }
                }
                In[p] = make_double2(nan_to_num(cu.x), nan_to_num(cu.y));
                *GP_ROW(Is, erow[p]) = In[p];
                par.Ia[p].x += In[p].x;
                par.Ia[p].y += In[p].y;
            }
            // update_voltage_backward (:207-232): phases of the child bus, masked by the parent
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if (brow[p] < 0 || prow[p] < 0) continue;
                double2 v = cur.Vb[p];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (erow[q] < 0) continue;
                    const double2 z = __ldg(&st->z[3 * p + q]);
                    v.x += z.x * In[q].x - z.y * In[q].y;
                    v.y += z.x * In[q].y + z.y * In[q].x;
                }
                par.Vb[p] = v;
            }
            cur = par;
        }
        // ---- root: final voltage and convergence (:123-128) ---------------------------
        ++it;
        diff = 0.0;
        bool nan = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            *GP_ROW(Vs, k.root_vrow[p]) = cur.Vb[p];
            const double dx = cur.Vb[p].x - k.vref[p].x, dy = cur.Vb[p].y - k.vref[p].y;
            const double d = sqrt(fma(dx, dx, dy * dy));
            nan |= (d != d);
            diff = fmax(diff, d);
        }
        if (nan) diff = __longlong_as_double(0x7ff8000000000000LL);
        converged = diff <= tol;
    }
#undef GP_ROW
#undef GP_CROW
    iters_out[s] = it;
    diff_out[s] = diff;
    status_out[s] = converged ? GP_ST_CONVERGED : ((diff != diff || isinf(diff)) ? GP_ST_NONFINITE : GP_ST_MAXITER);
}


// ---------------------------------------------------------------------------------------
// Tree-walk kernel (fbs_kernel = 6; the automatic choice for feeders that do not fit on chip).
//
// Same arithmetic as fbs_solve_kernel in depth-first order (see there), but the walk carries the
// state of the bus it stands on in REGISTERS instead of passing every value through memory:
//   cv[3]  the voltage of the current bus: its forward value, overwritten phase by phase as the
//          child that survives the reference's overwrite order (the one earliest in topo order,
//          visited last) rewrites it - the two never coexist, which is why one array suffices;
//   ca[3]  the sum of the new currents of the children visited so far.
// An event is one of
//   DESC  descend into a bus that has children: forward step, the child becomes the current bus
//         (the parent's voltage is already in its V rows; its partial current sum is parked in
//         its own - at that moment unused - I rows);
//   LEAF  a bus without children: forward step, load current, parent rewrite and the
//         contribution to the parent's current sum, all from registers;
//   ASC   come back up from a bus whose children are done: its backward step from registers,
//         then the parent becomes the current bus again (its voltage is taken back from its V
//         rows except for the phases this bus has just rewritten, its parked current sum from
//         its I rows).
// Per iteration a thread therefore reads from memory only what it cannot know: last iteration's
// edge currents, the loads, and a parent's state when it returns from a subtree - none of which
// depends on the arithmetic in flight.  The rows of event e+1 that come from DRAM (edge currents,
// loads, parked sums) are requested while event e is computed, so a thread never sits in front of
// a DRAM round trip with nothing to do; V rows are only ever stored (and re-read after a subtree).
// Stores: every V row once per iteration with its forward value, plus the surviving rewrites;
// every owned I row once (plus parked partial sums at the few branch points).
//
// Values: those of the two-sweep kernels up to the order in which a bus's child currents are
// added (visit order instead of adjacency order: a rounding-level difference, <= 1e-15 relative);
// iteration counts and status are the same (tests/test_fbs_gpu.py, tests/test_fbs_walk_host.py).
// ---------------------------------------------------------------------------------------
struct __align__(16) WalkEv {
    int step;
    int flags;      // 0-1 type; 2 VR edge; 3-5 bus phases; 6-8 I rows this step stores; 9-11 parent phases it rewrites;
                    // 12-14 phases added to the parent's current sum; 15-17 phases with a load or capacitor;
                    // 18-20 edge phases; 21-23 compile-time mask of a uniform step (0 = runtime masks)
    int e_row[3];   // this step's I rows as addressed (shared rows resolved), -1 = none / always zero
    int ld_i[3];    // rows requested ahead into oI: DESC/LEAF last iteration's edge current, ASC the parent's parked sum
    int ld_v[3];    // ASC: the parent's V rows to take back (-1: phase rewritten by this step, or absent)
    int st_ca[3];   // DESC: rows to park the parent's partial current sum in (-1: nothing to park)
    int ld_s[3];    // LEAF/ASC: load rows of the bus (-1: none)
    int pad[3];
};
static_assert(sizeof(WalkEv) == 80, "WalkEv layout");
constexpr int WK_DESC = 0, WK_LEAF = 1, WK_ASC = 2;

// The record the kernel reads (64 bytes, four 16-byte words; built from WalkEv by fbs_pack_walk):
//   w0 = {step, flags, a0, a1}   w1 = {a2, b0, b1, b2}   w2 = {c0, c1, c2, d0}   w3 = {d1, d2, 0, 0}
//   a = e_row (the step's I rows), b = ld_s (load rows),
//   c = st_ca for DESC (rows to park the parent's current sum in; flag bit 24 = there are any),
//       ld_i for ASC (rows holding the parent's parked sum), d = ld_v for ASC (the parent's V rows to take back)
struct __align__(16) WalkRec { int4 w[4]; };

// ZIP load + capacitor current of one bus phase (calc_sV solution.py:177-220, update_current :256-288)
__device__ __forceinline__ void walk_add_load_current(double2& In, const double2 v, const double2 sp, const double cap,
                                                      const FbsConst& k) {
    const double m2 = fma(v.x, v.x, v.y * v.y);
    const double a = sqrt(m2);
    const double a2 = a * a;
    const double f = k.aPQ + k.aI * a + k.aZ * a2;
    const double sx = sp.x * f;
    const double sy = sp.y * f - cap * a2;
    if (v.x != 0.0 || v.y != 0.0) {
        const double inv = 1.0 / m2;
        In.x += (sx * v.x + sy * v.y) * inv;
        In.y -= (sy * v.x - sx * v.y) * inv;
    }
}

#define GP_ROWP(base, row) ((base) + (size_t)(unsigned)(row) * (size_t)ldb)
#define GP_LD2(base, row) __ldcg(reinterpret_cast<const double2*>(GP_ROWP(base, row)))
#define GP_ST2(base, row, v) (*reinterpret_cast<double2*>(GP_ROWP(base, row)) = (v))
// v -= z * i  /  v += z * i  (complex, the three-operation form of the other FBS kernels)
#define GP_CMSUB(v, z, i) { v.x -= z.x * i.x - z.y * i.y; v.y -= z.x * i.y + z.y * i.x; }
#define GP_CMADD(v, z, i) { v.x += z.x * i.x - z.y * i.y; v.y += z.x * i.y + z.y * i.x; }

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
fbs_walk_kernel(const FbsStep* __restrict__ steps, const WalkRec* __restrict__ events, int n_events, FbsConst k,
                int n_vrows, int n_irows, long long S, unsigned ldb, const char* __restrict__ spu, char* __restrict__ V,
                char* __restrict__ I, int* __restrict__ iters_out, int* __restrict__ status_out,
                double* __restrict__ diff_out, double tol, int maxiter, const int2* __restrict__ alias, int n_alias) {
    const long long s = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (s >= S) return;
    char* const Vb = V + s * 16;
    char* const Ib = I + s * 16;
    const char* const Sb = spu + s * 16;
    const double2 zero = make_double2(0.0, 0.0);
    int it = 0;
    double diff = -1.0;
    bool converged = false;
    while (!converged && it < maxiter) {
        const bool first = it == 0;     // flat start: the currents are zero and are not read
        // the current bus: voltage per phase and the sum of the currents of the children visited so far
        double2 cv0 = k.vref[0], cv1 = k.vref[1], cv2 = k.vref[2], ca0 = zero, ca1 = zero, ca2 = zero;
        GP_ST2(Vb, k.root_vrow[0], cv0);                    // V[root] = Vref (:89)
        GP_ST2(Vb, k.root_vrow[1], cv1);
        GP_ST2(Vb, k.root_vrow[2], cv2);
        // operands requested one event ahead: pI = last iteration's edge currents (DESC, LEAF) or the parent's
        // parked current sum (ASC); pS = the loads of the bus (LEAF, ASC)
        double2 pI0 = zero, pI1 = zero, pI2 = zero, pS0 = zero, pS1 = zero, pS2 = zero;
        const int4* ev4 = reinterpret_cast<const int4*>(events);
        int4 n0 = __ldg(ev4), n1 = __ldg(ev4 + 1);
#define GP_WALK_REQUEST(q)                                                                      \
        {                                                                                       \
            int i0 = n0.z, i1 = n0.w, i2 = n1.x;                                                \
            if ((n0.y & 3) == WK_ASC) {                                                         \
                const int4 c = __ldg((q) + 2);                                                  \
                i0 = c.x; i1 = c.y; i2 = c.z;                                                   \
            } else if (first) {                                                                 \
                i0 = i1 = i2 = -1;                                                              \
            }                                                                                   \
            pI0 = i0 >= 0 ? GP_LD2(Ib, i0) : zero;                                              \
            pI1 = i1 >= 0 ? GP_LD2(Ib, i1) : zero;                                              \
            pI2 = i2 >= 0 ? GP_LD2(Ib, i2) : zero;                                              \
            pS0 = n1.y >= 0 ? GP_LD2(Sb, n1.y) : zero;                                          \
            pS1 = n1.z >= 0 ? GP_LD2(Sb, n1.z) : zero;                                          \
            pS2 = n1.w >= 0 ? GP_LD2(Sb, n1.w) : zero;                                          \
        }
        GP_WALK_REQUEST(ev4)
        for (int e = 0; e < n_events; ++e, ev4 += 4) {
            const int4 r0 = n0, r1 = n1;
            const double2 oI0 = pI0, oI1 = pI1, oI2 = pI2, oS0 = pS0, oS1 = pS1, oS2 = pS2;
            if (e + 1 < n_events) {
                n0 = __ldg(ev4 + 4);
                n1 = __ldg(ev4 + 5);
                GP_WALK_REQUEST(ev4 + 4)
            }
            const int fl = r0.y, type = fl & 3;
            const FbsStep* __restrict__ st = steps + r0.x;
            const int mb = (fl >> 3) & 7, me = (fl >> 18) & 7, mw = (fl >> 9) & 7;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_vrow));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_vrow));
            double2 rV0 = zero, rV1 = zero, rV2 = zero, pk0 = zero, pk1 = zero, pk2 = zero;
            if (type == WK_ASC) {
                // the parent's V rows and its parked current sum come back at the end of this event: requested now
                const int4 c = __ldg(ev4 + 2), d = __ldg(ev4 + 3);
                if (c.w >= 0) rV0 = GP_LD2(Vb, c.w);
                if (d.x >= 0) rV1 = GP_LD2(Vb, d.x);
                if (d.y >= 0) rV2 = GP_LD2(Vb, d.y);
                pk0 = oI0; pk1 = oI1; pk2 = oI2;
            } else if (fl & (1 << 24)) {    // DESC at a branch point: park the parent's partial current sum in its own I rows
                const int4 c = __ldg(ev4 + 2);
                if (c.x >= 0) GP_ST2(Ib, c.x, ca0);
                if (c.y >= 0) GP_ST2(Ib, c.y, ca1);
                if (c.z >= 0) GP_ST2(Ib, c.z, ca2);
            }
            double2 b0 = cv0, b1 = cv1, b2 = cv2;           // this bus's voltage
            double2 In0 = zero, In1 = zero, In2 = zero;     // its edge current
            if (fl & 4) {
                // regulator edge: vr_forward :138-160 on the way down, vr_backward :186-205 on the way up, no current
                if (type != WK_ASC) {
                    const double g0 = __ldg(&st->gamma[0]), g1 = __ldg(&st->gamma[1]), g2 = __ldg(&st->gamma[2]);
                    b0 = (mb & 1) ? make_double2(g0 * cv0.x, g0 * cv0.y) : zero;
                    b1 = (mb & 2) ? make_double2(g1 * cv1.x, g1 * cv1.y) : zero;
                    b2 = (mb & 4) ? make_double2(g2 * cv2.x, g2 * cv2.y) : zero;
                    if (mb & 1) GP_ST2(Vb, bv.x, b0);
                    if (mb & 2) GP_ST2(Vb, bv.y, b1);
                    if (mb & 4) GP_ST2(Vb, bv.z, b2);
                }
                if (type == WK_DESC) {
                    cv0 = b0; cv1 = b1; cv2 = b2;
                    ca0 = ca1 = ca2 = zero;
                    continue;
                }
                double2 w0 = zero, w1 = zero, w2 = zero;
                if (mw & 1) { const double ig = __ldg(&st->inv_gamma[0]); w0 = make_double2(ig * b0.x, ig * b0.y); GP_ST2(Vb, pv.x, w0); }
                if (mw & 2) { const double ig = __ldg(&st->inv_gamma[1]); w1 = make_double2(ig * b1.x, ig * b1.y); GP_ST2(Vb, pv.y, w1); }
                if (mw & 4) { const double ig = __ldg(&st->inv_gamma[2]); w2 = make_double2(ig * b2.x, ig * b2.y); GP_ST2(Vb, pv.z, w2); }
                if (type == WK_ASC) {
                    cv0 = (mw & 1) ? w0 : rV0; cv1 = (mw & 2) ? w1 : rV1; cv2 = (mw & 4) ? w2 : rV2;
                    ca0 = pk0; ca1 = pk1; ca2 = pk2;
                } else {
                    if (mw & 1) cv0 = w0;
                    if (mw & 2) cv1 = w1;
                    if (mw & 4) cv2 = w2;
                }
                continue;
            }
            if (type != WK_ASC) {
                // ---- forward step on the way down (update_voltage_forward :162-184) ----
                if (!(mb & 1)) b0 = zero;
                if (!(mb & 2)) b1 = zero;
                if (!(mb & 4)) b2 = zero;
                if (!first) {
                    if (me & 1) {
                        if (mb & 1) { const double2 z = __ldg(&st->z[0]); GP_CMSUB(b0, z, oI0) }
                        if (mb & 2) { const double2 z = __ldg(&st->z[3]); GP_CMSUB(b1, z, oI0) }
                        if (mb & 4) { const double2 z = __ldg(&st->z[6]); GP_CMSUB(b2, z, oI0) }
                    }
                    if (me & 2) {
                        if (mb & 1) { const double2 z = __ldg(&st->z[1]); GP_CMSUB(b0, z, oI1) }
                        if (mb & 2) { const double2 z = __ldg(&st->z[4]); GP_CMSUB(b1, z, oI1) }
                        if (mb & 4) { const double2 z = __ldg(&st->z[7]); GP_CMSUB(b2, z, oI1) }
                    }
                    if (me & 4) {
                        if (mb & 1) { const double2 z = __ldg(&st->z[2]); GP_CMSUB(b0, z, oI2) }
                        if (mb & 2) { const double2 z = __ldg(&st->z[5]); GP_CMSUB(b1, z, oI2) }
                        if (mb & 4) { const double2 z = __ldg(&st->z[8]); GP_CMSUB(b2, z, oI2) }
                    }
                }
                if (mb & 1) GP_ST2(Vb, bv.x, b0);
                if (mb & 2) GP_ST2(Vb, bv.y, b1);
                if (mb & 4) GP_ST2(Vb, bv.z, b2);
                if (type == WK_DESC) {
                    cv0 = b0; cv1 = b1; cv2 = b2;
                    ca0 = ca1 = ca2 = zero;
                    continue;
                }
            } else {
                In0 = ca0; In1 = ca1; In2 = ca2;            // the children's currents (:274-284)
            }
            // ---- backward step of a leaf, or on the way up: calc_sV, update_current, update_voltage_backward ----
            {
                const int ml = (fl >> 15) & 7;
                // a phase without load and capacitor adds conj(0 / V) = 0 - unless V is not finite, when the
                // reference's 0 * inf makes the current NaN (nan_to_num then zeroes it): x - x does both
                if (me & 1) {
                    if (ml & 1) walk_add_load_current(In0, b0, oS0, __ldg(&st->cap[0]), k);
                    else { const double t = (b0.x - b0.x) + (b0.y - b0.y); In0.x += t; In0.y += t; }
                }
                if (me & 2) {
                    if (ml & 2) walk_add_load_current(In1, b1, oS1, __ldg(&st->cap[1]), k);
                    else { const double t = (b1.x - b1.x) + (b1.y - b1.y); In1.x += t; In1.y += t; }
                }
                if (me & 4) {
                    if (ml & 4) walk_add_load_current(In2, b2, oS2, __ldg(&st->cap[2]), k);
                    else { const double t = (b2.x - b2.x) + (b2.y - b2.y); In2.x += t; In2.y += t; }
                }
                double sum = 0.0;
                if (me & 1) sum += fabs(In0.x) + fabs(In0.y);
                if (me & 2) sum += fabs(In1.x) + fabs(In1.y);
                if (me & 4) sum += fabs(In2.x) + fabs(In2.y);
                if (!(sum <= 1.7976931348623157e308)) {         // nan_to_num (:288) acts on non-finite values only
                    if (me & 1) In0 = make_double2(nan_to_num(In0.x), nan_to_num(In0.y));
                    if (me & 2) In1 = make_double2(nan_to_num(In1.x), nan_to_num(In1.y));
                    if (me & 4) In2 = make_double2(nan_to_num(In2.x), nan_to_num(In2.y));
                }
                if (!(me & 1)) In0 = zero;
                if (!(me & 2)) In1 = zero;
                if (!(me & 4)) In2 = zero;
                const int ms = (fl >> 6) & 7;
                if (ms & 1) GP_ST2(Ib, r0.z, In0);
                if (ms & 2) GP_ST2(Ib, r0.w, In1);
                if (ms & 4) GP_ST2(Ib, r1.x, In2);
            }
            // the parent phases this bus rewrites (the surviving stores of update_voltage_backward :207-232)
            double2 w0 = b0, w1 = b1, w2 = b2;
            if (mw) {
                if (me & 1) {
                    if (mw & 1) { const double2 z = __ldg(&st->z[0]); GP_CMADD(w0, z, In0) }
                    if (mw & 2) { const double2 z = __ldg(&st->z[3]); GP_CMADD(w1, z, In0) }
                    if (mw & 4) { const double2 z = __ldg(&st->z[6]); GP_CMADD(w2, z, In0) }
                }
                if (me & 2) {
                    if (mw & 1) { const double2 z = __ldg(&st->z[1]); GP_CMADD(w0, z, In1) }
                    if (mw & 2) { const double2 z = __ldg(&st->z[4]); GP_CMADD(w1, z, In1) }
                    if (mw & 4) { const double2 z = __ldg(&st->z[7]); GP_CMADD(w2, z, In1) }
                }
                if (me & 4) {
                    if (mw & 1) { const double2 z = __ldg(&st->z[2]); GP_CMADD(w0, z, In2) }
                    if (mw & 2) { const double2 z = __ldg(&st->z[5]); GP_CMADD(w1, z, In2) }
                    if (mw & 4) { const double2 z = __ldg(&st->z[8]); GP_CMADD(w2, z, In2) }
                }
                if (mw & 1) GP_ST2(Vb, pv.x, w0);
                if (mw & 2) GP_ST2(Vb, pv.y, w1);
                if (mw & 4) GP_ST2(Vb, pv.z, w2);
            }
            if (type == WK_ASC) {       // the parent is the current bus again
                cv0 = (mw & 1) ? w0 : rV0; cv1 = (mw & 2) ? w1 : rV1; cv2 = (mw & 4) ? w2 : rV2;
                ca0 = pk0; ca1 = pk1; ca2 = pk2;
            } else {
                if (mw & 1) cv0 = w0;
                if (mw & 2) cv1 = w1;
                if (mw & 4) cv2 = w2;
            }
            const int mc = (fl >> 12) & 7;
            if (mc & 1) { ca0.x += In0.x; ca0.y += In0.y; }
            if (mc & 2) { ca1.x += In1.x; ca1.y += In1.y; }
            if (mc & 4) { ca2.x += In2.x; ca2.y += In2.y; }
        }
#undef GP_WALK_REQUEST
        ++it;
        // root mismatch (:123-128) from the registers: the root is the current bus after the last event
        {
            const double dx0 = cv0.x - k.vref[0].x, dy0 = cv0.y - k.vref[0].y;
            const double dx1 = cv1.x - k.vref[1].x, dy1 = cv1.y - k.vref[1].y;
            const double dx2 = cv2.x - k.vref[2].x, dy2 = cv2.y - k.vref[2].y;
            const double d0 = sqrt(fma(dx0, dx0, dy0 * dy0)), d1 = sqrt(fma(dx1, dx1, dy1 * dy1)),
                         d2 = sqrt(fma(dx2, dx2, dy2 * dy2));
            diff = fmax(fmax(fmax(0.0, d0), d1), d2);
            if (d0 != d0 || d1 != d1 || d2 != d2) diff = __longlong_as_double(0x7ff8000000000000LL);
        }
        converged = diff <= tol;
    }
    for (int j = 0; j < n_alias; ++j) {         // shared I rows: the caller's array gets every row, once
        const int2 a = __ldg(alias + j);
        GP_ST2(Ib, a.x, a.y >= 0 ? GP_LD2(Ib, a.y) : zero);
    }
    iters_out[s] = it;
    diff_out[s] = diff;
    status_out[s] = fbs_status(converged, diff);
}
#undef GP_ROWP
#undef GP_LD2
#undef GP_ST2
#undef GP_CMSUB
#undef GP_CMADD

// Derived results: sV, |V| per V row; Stx/Srx per line row; loss per scenario.
__global__ void fbs_post_rows_kernel(int n_vrows, const int* __restrict__ vrow_srow,
                                     const double* __restrict__ vrow_cap, FbsConst k, long long S,
                                     long long ld, const double2* __restrict__ spu,
                                     const double2* __restrict__ V, double2* __restrict__ sV,
                                     double* __restrict__ Vmag, const PeerDst peers,
                                     const double* __restrict__ wpu, const int* __restrict__ vv) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int r = blockIdx.y; r < n_vrows; r += gridDim.y) {
        const double2 v = V[(long long)r * ld + s];
        double a;
        const double a2 = cabs2(v, a);
        if (Vmag) Vmag[(long long)r * ld + s] = a;
#pragma unroll
        for (int q = 0; q < GP_MAX_PEERS; ++q)      // fused all-gather: the same value to every rank's buffer
            if (q < peers.n) peers.vmag[q][(long long)r * ld + s] = a;
        if (sV) {
            double2 sv = make_double2(0.0, 0.0);
            const int lr = __ldg(vrow_srow + r);
            if (lr >= 0) {
                const double2 sp = spu[(long long)lr * ld + s];
                const double f = k.aPQ + k.aI * a + k.aZ * a2;
                sv.x = sp.x * f;
                sv.y = sp.y * f;
            }
            sv.y -= __ldg(vrow_cap + r) * a2;
            if (wpu || vv) {                        // + 1j * wpu: the caller's row, or the droop of the row's controller
                double w = wpu ? wpu[(long long)r * ld + s] : 0.0;
                if (vv && __ldg(vv + r)) w = fbs_droop(a, k);
                sv.y += w;
            }
            sV[(long long)r * ld + s] = sv;
        }
    }
}

__global__ void fbs_post_lines_kernel(int n_lines, const int* __restrict__ lines, long long S,
                                      long long ld, const double2* __restrict__ V,
                                      const double2* __restrict__ I, double2* __restrict__ Stx,
                                      double2* __restrict__ Srx, double2* __restrict__ loss, const PeerDst peers) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double2 acc = make_double2(0.0, 0.0);
    for (int l = 0; l < n_lines; ++l) {
        const int* rec = lines + 9 * l;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int ir = __ldg(rec + p);
            if (ir < 0) continue;   // I is 0 on phases the line lacks -> S = 0
            const int tr = __ldg(rec + 3 + p), rr = __ldg(rec + 6 + p);
            const double2 i = I[(long long)ir * ld + s];
            const double2 vt = tr >= 0 ? V[(long long)tr * ld + s] : make_double2(0.0, 0.0);
            const double2 vr = rr >= 0 ? V[(long long)rr * ld + s] : make_double2(0.0, 0.0);
            // V * conj(I)
            const double2 st = make_double2(vt.x * i.x + vt.y * i.y, vt.y * i.x - vt.x * i.y);
            const double2 sr = make_double2(vr.x * i.x + vr.y * i.y, vr.y * i.x - vr.x * i.y);
            if (Stx) Stx[(long long)ir * ld + s] = st;
            if (Srx) Srx[(long long)ir * ld + s] = sr;
            acc.x += st.x - sr.x;
            acc.y += st.y - sr.y;
        }
    }
    if (loss) loss[s] = acc;
#pragma unroll
    for (int q = 0; q < GP_MAX_PEERS; ++q)
        if (q < peers.n) peers.loss[q][s] = acc;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_version(void) { return 100; }

extern "C" int gp_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return GP_EINVAL;
    strncpy(buf, g_err, len - 1);
    buf[len - 1] = 0;
    return GP_OK;
}

extern "C" int64_t gp_launch_count(void) { return g_launches.load(); }

extern "C" int gp_set_option(const char* name, int value) {
    if (name && strcmp(name, "peer_push_streams") == 0 && value >= 1 && value <= 8) {
        g_peer_push_streams = value;
        return GP_OK;
    }
    if (name && strcmp(name, "fbs_auto_walk") == 0 && (value == 0 || value == 1)) {
        g_fbs_auto_walk = value;
        return GP_OK;
    }
    if (name && strcmp(name, "fbs_walk_blocks") == 0 && (value == 4 || value == 5 || value == 6 || value == 8)) {
        g_fbs_walk_blocks = value;
        return GP_OK;
    }
    if (name && strcmp(name, "fbs_kernel") == 0 && value >= 0 && value <= 6) {
        g_fbs_kernel = value;
        return GP_OK;
    }
    if (name && strcmp(name, "fbs_prefetch") == 0 && ((value >= -1 && value <= 64) || (value >= 100 && value <= 164))) {
        g_fbs_prefetch = value;
        return GP_OK;
    }
    if (name && strcmp(name, "fbs_share_rows") == 0 && (value == 0 || value == 1)) {
        g_fbs_share_rows = value;
        return GP_OK;
    }
    if (name && strcmp(name, "fbs_order") == 0 && (value == 0 || value == 1)) {
        g_fbs_order = value;
        return GP_OK;
    }
    if (name && strcmp(name, "nr3_ring") == 0 && value >= -1 && value <= 1) {
        g_nr3_ring = value;
        return GP_OK;
    }
    if (name && strcmp(name, "nr3_kernel") == 0 && value >= 0 && value <= 2) {
        g_nr3_kernel = value;
        return GP_OK;
    }
    if (name && strcmp(name, "nr3_prefetch") == 0 && value >= -1 && value <= 64) {
        g_nr3_prefetch = value;
        return GP_OK;
    }
    if (name && strcmp(name, "jit_dfs") == 0 && value >= -1 && value <= 1) {
        g_jit_dfs = value;          // takes effect for feeders created afterwards
        return GP_OK;
    }
    if (name && strcmp(name, "jit_i_global") == 0 && value >= -1 && value <= 1) {
        g_jit_i_global = value;      // takes effect for feeders created afterwards
        return GP_OK;
    }
    if (name && strcmp(name, "host_zero_copy") == 0 && (value == 0 || value == 1)) {
        g_host_zero_copy = value;
        return GP_OK;
    }
    if (name && strcmp(name, "host_read_ahead") == 0 && (value == 0 || value == 1)) {
        g_host_read_ahead = value;
        return GP_OK;
    }
    if (name && strcmp(name, "host_warps") == 0 && value >= 1 && value <= 8) {
        g_host_warps = value;
        return GP_OK;
    }
    if (name && strcmp(name, "host_chunks") == 0 && value >= 0 && value <= 64) {
        g_host_chunks = value;
        return GP_OK;
    }
    return fail(GP_EINVAL, "gp_set_option: unknown option");
}

extern "C" int gp_fbs_onchip_warps(GpFeeder* f) {
    if (!f || f->kind != 1) return fail(GP_EINVAL, "gp_fbs_onchip_warps: not an FBS feeder handle");
    return ((FbsImpl*)f->impl)->onchip_warps;
}

extern "C" int gp_fbs_jit_warps(GpFeeder* f) {
    if (!f || f->kind != 1) return fail(GP_EINVAL, "gp_fbs_jit_warps: not an FBS feeder handle");
    return ((FbsImpl*)f->impl)->jit_plan_warps;
}

extern "C" int gp_fbs_active_kernel(GpFeeder* f) {
    if (!f || f->kind != 1) return fail(GP_EINVAL, "gp_fbs_active_kernel: not an FBS feeder handle");
    const FbsImpl* im = (const FbsImpl*)f->impl;
    if (g_fbs_kernel == 5 || (g_fbs_kernel == 0 && im->jit_kernel)) return 5;
    if ((g_fbs_kernel == 6 || (g_fbs_kernel == 0 && g_fbs_auto_walk)) && im->n_walk > 0) return 6;
    if (g_fbs_kernel == 4) return 4;
    if (g_fbs_kernel == 2) return 2;
    if (g_fbs_kernel == 1 && im->n_events > 0 && im->dfs_stack_ok) return 1;
    return 3;
}

namespace gp {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (double)(threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], b, c);
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += a[j];
    if (sum == 12345.678) out[0] = sum;    // never true: keeps the chains alive
}
}  // namespace gp

extern "C" int gp_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(GP_EINVAL, "gp_measure_fp64_peak: null argument");
    GP_CUDA(cudaSetDevice(device));
    int n_sm = 0;
    GP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    double* d_out = nullptr;
    GP_CUDA(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1;
    GP_CUDA(cudaEventCreate(&e0));
    GP_CUDA(cudaEventCreate(&e1));
    const int iters = 8192, blocks = n_sm * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        GP_CUDA(cudaEventRecord(e0, 0));
        fp64_peak_kernel<<<blocks, 256>>>(d_out, iters, 0.999999, 1e-9);
        GP_CUDA(cudaEventRecord(e1, 0));
        GP_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        GP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    GP_CUDA(cudaGetLastError());
    *tflops = best;
    return GP_OK;
}

// ---- FP64 tensor-core (DMMA) probe ---------------------------------------------------------------------
// mode 0: 8 independent chains of mma.sync.m8n8k4.f64 per warp; mode 1: 8 chains of DFMA per thread; mode 2: both
// interleaved in one instruction stream.  If DMMA ran on its own pipe next to the FP64 FMA pipe, mode 2 would take
// about max(t0, t1); if they share the pipe (or the issue slots), about t0 + t1.
namespace gp {
template <int MODE>
__global__ void __launch_bounds__(256) fp64_dmma_probe_kernel(double* out, int iters, double b, double c) {
    double acc[8][2], f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j][0] = (double)(threadIdx.x + j);
        acc[j][1] = (double)(threadIdx.x - j);
        f[j] = (double)(threadIdx.x + 3 * j);
    }
    const double a_frag = b, b_frag = c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0 || MODE == 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(acc[j][0]), "+d"(acc[j][1]) : "d"(a_frag), "d"(b_frag));
            if (MODE == 1 || MODE == 2) f[j] = fma(f[j], b, c);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += acc[j][0] + acc[j][1] + f[j];
    if (sum == 12345.678) out[0] = sum;
}
}  // namespace gp

// out[0] = DMMA TFLOP/s (512 flop per warp instruction), out[1] = DFMA TFLOP/s (this kernel's 8 chains),
// out[2], out[3] = the same two rates while both run interleaved, out[4..6] = the three durations in ms
extern "C" int gp_measure_dmma(int device, double* out7) {
    if (!out7) return fail(GP_EINVAL, "gp_measure_dmma: null argument");
    GP_CUDA(cudaSetDevice(device));
    int n_sm = 0;
    GP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    double* d_out = nullptr;
    GP_CUDA(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1;
    GP_CUDA(cudaEventCreate(&e0));
    GP_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = n_sm * 8;
    double ms[3] = {1e30, 1e30, 1e30};
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 5; ++rep) {
            GP_CUDA(cudaEventRecord(e0, 0));
            if (mode == 0) fp64_dmma_probe_kernel<0><<<blocks, 256>>>(d_out, iters, 0.999999, 1e-9);
            else if (mode == 1) fp64_dmma_probe_kernel<1><<<blocks, 256>>>(d_out, iters, 0.999999, 1e-9);
            else fp64_dmma_probe_kernel<2><<<blocks, 256>>>(d_out, iters, 0.999999, 1e-9);
            GP_CUDA(cudaEventRecord(e1, 0));
            GP_CUDA(cudaEventSynchronize(e1));
            float t = 0.f;
            GP_CUDA(cudaEventElapsedTime(&t, e0, e1));
            if (rep > 0 && t < ms[mode]) ms[mode] = t;
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    GP_CUDA(cudaGetLastError());
    const double warps = (double)blocks * 8.0, threads = (double)blocks * 256.0;
    const double mma_flop = 512.0 * 8.0 * iters * warps, fma_flop = 2.0 * 8.0 * iters * threads;
    out7[0] = mma_flop / (ms[0] * 1e-3) / 1e12;
    out7[1] = fma_flop / (ms[1] * 1e-3) / 1e12;
    out7[2] = mma_flop / (ms[2] * 1e-3) / 1e12;
    out7[3] = fma_flop / (ms[2] * 1e-3) / 1e12;
    out7[4] = ms[0];
    out7[5] = ms[1];
    out7[6] = ms[2];
    return GP_OK;
}

// GpFbsDesc -> step records, child table, constants (host side; shared by gp_fbs_create and
// the source generator)
static int fbs_tables_from_desc(const GpFbsDesc* d, std::vector<FbsStep>& steps, std::vector<int4>& child4,
                                FbsConst& k, std::vector<int>& vrow_srow, std::vector<double>& vrow_cap) {
    if (d->n_steps < 0 || d->n_vrows <= 0 || d->n_irows < 0 || d->n_srows < 0 || d->n_child < 0 || d->n_lines < 0)
        return fail(GP_EINVAL, "gp_fbs_create: bad sizes");
    for (int l = 0; l < 3 * d->n_lines; ++l)
        if (d->line_irow[l] < -1 || d->line_irow[l] >= d->n_irows || d->line_tx_vrow[l] < -1 ||
            d->line_tx_vrow[l] >= d->n_vrows || d->line_rx_vrow[l] < -1 || d->line_rx_vrow[l] >= d->n_vrows)
            return fail(GP_EINVAL, "gp_fbs_create: line %d row index out of range", l / 3);
    steps.assign((size_t)d->n_steps, FbsStep());
    vrow_srow.assign((size_t)d->n_vrows, -1);
    vrow_cap.assign((size_t)d->n_vrows, 0.0);
    auto chk = [&](int r, int n) { return r >= -1 && r < n; };
    for (int i = 0; i < d->n_steps; ++i) {
        FbsStep& s = steps[i];
        memset(&s, 0, sizeof(s));
        s.kind = d->kind[i];
        s.child_begin = d->child_begin[i];
        s.child_count = d->child_count[i];
        if (s.child_begin < 0 || s.child_count < 0 || s.child_begin + s.child_count > d->n_child)
            return fail(GP_EINVAL, "gp_fbs_create: step %d child range out of bounds", i);
        for (int p = 0; p < 3; ++p) {
            s.bus_vrow[p] = d->bus_vrow[3 * i + p];
            s.par_vrow[p] = d->par_vrow[3 * i + p];
            s.edge_irow[p] = d->edge_irow[3 * i + p];
            s.load_srow[p] = d->load_srow[3 * i + p];
            if (!chk(s.bus_vrow[p], d->n_vrows) || !chk(s.par_vrow[p], d->n_vrows) ||
                !chk(s.edge_irow[p], d->n_irows) || !chk(s.load_srow[p], d->n_srows))
                return fail(GP_EINVAL, "gp_fbs_create: step %d row index out of range", i);
            s.cap[p] = d->cappu[3 * i + p];
            s.gamma[p] = d->vr_gamma[3 * i + p];
            s.inv_gamma[p] = s.gamma[p] != 0.0 ? 1.0 / s.gamma[p] : 0.0;
            if (s.bus_vrow[p] >= 0) {
                vrow_srow[s.bus_vrow[p]] = s.load_srow[p];
                vrow_cap[s.bus_vrow[p]] = s.cap[p];
            }
            for (int q = 0; q < 3; ++q) {
                s.z[3 * p + q].x = d->z[((size_t)i * 9 + 3 * p + q) * 2];
                s.z[3 * p + q].y = d->z[((size_t)i * 9 + 3 * p + q) * 2 + 1];
            }
        }
    }
    // rows of a per-scenario gamma array: regulated (step, phase) pairs in step, then phase order
    {
        int g = 0;
        for (int i = 0; i < d->n_steps; ++i)
            for (int p = 0; p < 3; ++p)
                steps[i].gam_row[p] = (steps[i].kind == GP_EDGE_VR && steps[i].gamma[p] != 0.0) ? g++ : -1;
    }
    // fast mask (fwd_fast / bwd_fast): a Line/Transformer edge whose phases are exactly the bus
    // phases, all present on the parent.  Stored above the kind byte.
    for (int i = 0; i < d->n_steps; ++i) {
        FbsStep& s = steps[i];
        if (s.kind != GP_EDGE_LINE) continue;
        int mb = 0, me = 0, mp = 0;
        for (int p = 0; p < 3; ++p) {
            mb |= (s.bus_vrow[p] >= 0) << p;
            me |= (s.edge_irow[p] >= 0) << p;
            mp |= (s.par_vrow[p] >= 0) << p;
        }
        if (mb != 0 && mb == me && (mp & mb) == mb) s.kind |= mb << 8;
    }
    // Dead parent-voltage stores (bits 19..21 of kind).  The backward sweep runs over reversed(topo_order) and
    // every child rewrites the phases it shares with its parent (solution_fbs.py:207-232): of the children of a
    // bus only the one EARLIEST in topo order survives per phase, nothing reads the others' values.  They are
    // not stored (IEEE 123: about 40 % of the backward sweep's V stores), which is also what lets the kernel
    // walk the tree depth-first on one V array.
    // bits 22..24: bus phases without load and capacitor - their load current is an exact zero and is not computed
    for (int i = 0; i < d->n_steps; ++i)
        for (int p = 0; p < 3; ++p)
            if (steps[i].bus_vrow[p] >= 0 && steps[i].load_srow[p] < 0 && steps[i].cap[p] == 0.0) steps[i].kind |= 1 << (22 + p);
    {
        std::vector<char> written((size_t)d->n_vrows, 0);
        for (int i = 0; i < d->n_steps; ++i) {
            FbsStep& s = steps[i];
            const bool vr = (s.kind & 0xff) == GP_EDGE_VR;
            for (int p = 0; p < 3; ++p) {
                const bool stores = s.par_vrow[p] >= 0 && (vr ? s.gamma[p] != 0.0 : s.bus_vrow[p] >= 0);
                if (!stores) continue;
                if (written[s.par_vrow[p]]) s.kind |= 1 << (19 + p);
                else written[s.par_vrow[p]] = 1;
            }
        }
    }
    for (int p = 0; p < 3; ++p) {
        k.root_vrow[p] = d->root_vrow[p];
        k.root_load_srow[p] = d->root_load_srow[p];
        k.root_cap[p] = d->root_cappu[p];
        k.vref[p] = make_double2(d->vref[2 * p], d->vref[2 * p + 1]);
        if (d->root_vrow[p] < 0 || d->root_vrow[p] >= d->n_vrows)
            return fail(GP_EINVAL, "gp_fbs_create: the root bus must carry all three phases");
        vrow_srow[d->root_vrow[p]] = d->root_load_srow[p];
        vrow_cap[d->root_vrow[p]] = d->root_cappu[p];
    }
    k.aPQ = d->zip[0];
    k.aI = d->zip[1];
    k.aZ = d->zip[2];
    child4.assign((size_t)d->n_child, make_int4(-1, -1, -1, 0));
    for (int c = 0; c < d->n_child; ++c) {
        child4[c] = make_int4(d->child_irow[3 * c], d->child_irow[3 * c + 1], d->child_irow[3 * c + 2], 0);
        if (!chk(child4[c].x, d->n_irows) || !chk(child4[c].y, d->n_irows) || !chk(child4[c].z, d->n_irows))
            return fail(GP_EINVAL, "gp_fbs_create: child %d row index out of range", c);
    }
    return GP_OK;
}

extern "C" int64_t gp_fbs_jit_source(const GpFbsDesc* d, int32_t warps, int32_t i_global, char* buf, int64_t cap) {
    if (!d || warps < 1 || warps > 16) return fail(GP_EINVAL, "gp_fbs_jit_source: bad argument");
    std::vector<FbsStep> steps;
    std::vector<int4> child4;
    std::vector<int> vrow_srow;
    std::vector<double> vrow_cap;
    FbsConst k;
    int rc = fbs_tables_from_desc(d, steps, child4, k, vrow_srow, vrow_cap);
    if (rc) return rc;
    std::vector<int> lines((size_t)std::max(d->n_lines, 0) * 9);
    for (int l = 0; l < d->n_lines; ++l)
        for (int p = 0; p < 3; ++p) {
            lines[9 * l + p] = d->line_irow[3 * l + p];
            lines[9 * l + 3 + p] = d->line_tx_vrow[3 * l + p];
            lines[9 * l + 6 + p] = d->line_rx_vrow[3 * l + p];
        }
    const std::string src = fbs_jit_source(steps, child4, lines, k, d->n_vrows, d->n_irows, d->n_srows, warps,
                                           (i_global & 1) != 0, false, (i_global & 2) != 0, (i_global & 4) != 0);
    if (src.empty())
        return fail(GP_EJIT, "gp_fbs_jit_source: non-finite impedance in the descriptor, or a depth-first build (mode 4) "
                             "combined with mode 1 / 2 or asked of a feeder with more than 64 steps");
    if (buf && cap > 0) {
        const size_t n = std::min((size_t)cap - 1, src.size());
        memcpy(buf, src.data(), n);
        buf[n] = 0;
    }
    return (int64_t)src.size();
}

extern "C" int gp_jit_compile(const char* source, void* cubin, int64_t cap, int64_t* cubin_bytes) {
    if (!source || !cubin_bytes) return fail(GP_EINVAL, "gp_jit_compile: null argument");
    std::vector<char> bin;
    std::string log;
    const int rc = jit_compile(source, bin, log);
    if (rc) return fail(rc, "gp_jit_compile: %s", log.substr(0, 400).c_str());
    *cubin_bytes = (int64_t)bin.size();
    if (cubin && cap >= (int64_t)bin.size()) memcpy(cubin, bin.data(), bin.size());
    return GP_OK;
}

// ---- driver API through dlopen (module load + launch of the specialised kernel) ------------
namespace {
struct Driver {
    void* h = nullptr;
    int (*ModuleLoadData)(void**, const void*) = nullptr;
    int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
    int (*ModuleUnload)(void*) = nullptr;
    int (*FuncSetAttribute)(void*, int, int) = nullptr;
    int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**,
                        void**) = nullptr;
    bool ok = false;
};
Driver& driver() {
    static Driver d;
    static bool tried = false;
    if (tried) return d;
    tried = true;
    d.h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!d.h) return d;
    *(void**)(&d.ModuleLoadData) = dlsym(d.h, "cuModuleLoadData");
    *(void**)(&d.ModuleGetFunction) = dlsym(d.h, "cuModuleGetFunction");
    *(void**)(&d.ModuleUnload) = dlsym(d.h, "cuModuleUnload");
    *(void**)(&d.FuncSetAttribute) = dlsym(d.h, "cuFuncSetAttribute");
    *(void**)(&d.LaunchKernel) = dlsym(d.h, "cuLaunchKernel");
    d.ok = d.ModuleLoadData && d.ModuleGetFunction && d.ModuleUnload && d.FuncSetAttribute && d.LaunchKernel;
    return d;
}
}  // namespace

// Generate, compile and load one variant of the feeder-specialised kernel: the device-buffer kernel
// (read_ahead = false) or the kernel of the one-launch host path, which fetches a warp's next tile over
// PCIe before it stores the current one (read_ahead = true; more registers, so it is its own build).
static int fbs_jit_build(GpFeeder* f, int nw, bool ig, bool read_ahead, void** mod_out, void** fn_out, bool ex = false,
                         bool dfs = false) {
    FbsImpl* im = (FbsImpl*)f->impl;
    GP_CUDA(cudaSetDevice(f->device));
    GP_CUDA(cudaFree(0));       // make sure the primary context exists and is current
    const std::string src = fbs_jit_source(im->h_steps, im->h_child, im->h_lines, im->k, im->n_vrows, im->n_irows,
                                           im->n_srows, nw, ig, read_ahead, ex, dfs);
    if (src.empty()) return fail(GP_EJIT, "gp_fbs_specialise: non-finite impedance");
    std::vector<char> bin;
    std::string log;
    int rc = jit_compile(src, bin, log);
    if (rc) return fail(rc, "gp_fbs_specialise: %s", log.substr(0, 400).c_str());
    Driver& dr = driver();
    if (!dr.ok) return fail(GP_EJIT, "gp_fbs_specialise: libcuda.so.1 could not be loaded");
    void* mod = nullptr;
    void* fn = nullptr;
    int cr = dr.ModuleLoadData(&mod, bin.data());
    if (cr != 0) return fail(GP_EJIT, "gp_fbs_specialise: cuModuleLoadData failed (%d)", cr);
    cr = dr.ModuleGetFunction(&fn, mod, "gp_fbs_jit_kernel");
    const int rows = dfs ? im->n_islots + im->n_dfs_stack : im->n_vrows + (ig ? 0 : im->n_islots);
    const int smem = nw * rows * 512;
    if (cr == 0) cr = dr.FuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, smem);
    if (cr != 0) {
        dr.ModuleUnload(mod);
        return fail(GP_EJIT, "gp_fbs_specialise: kernel lookup / shared-memory opt-in failed (%d)", cr);
    }
    *mod_out = mod;
    *fn_out = fn;
    return GP_OK;
}

extern "C" int gp_fbs_specialise(GpFeeder* f) {
    if (!f || f->kind != 1) return fail(GP_EINVAL, "gp_fbs_specialise: not an FBS feeder handle");
    FbsImpl* im = (FbsImpl*)f->impl;
    if (im->jit_kernel) return GP_OK;
    if (im->jit_plan_warps < 1) return fail(GP_EJIT, "gp_fbs_specialise: the feeder does not fit the on-chip kernel");
    void *mod = nullptr, *fn = nullptr;
    int rc = fbs_jit_build(f, im->jit_plan_warps, im->jit_plan_ig, false, &mod, &fn, false, im->jit_plan_dfs);
    if (rc) return rc;
    im->jit_warps = im->jit_plan_warps;
    im->jit_ig = im->jit_plan_ig;
    im->jit_dfs = im->jit_plan_dfs;
    im->jit_rows = im->jit_dfs ? im->n_islots + im->n_dfs_stack : im->n_vrows + (im->jit_ig ? 0 : im->n_islots);
    im->jit_module = mod;
    im->jit_kernel = fn;
    return GP_OK;
}

// Depth-first schedule of the feeder: parent step of every step, children visited latest-in-topo-order first,
// one descend and one ascend event {step, (level << 4) | flags} per step.  The streaming kernel runs its forward
// and backward steps in this order (fbs_order); fbs_dfs_kernel additionally keeps the path state on a stack, for
// which regulators must cover every phase of their bus and the tree must be at most 64 deep (stack_ok).
// Empty when the step table is not a tree in topo order.
static void fbs_build_events(const GpFbsDesc* d, const std::vector<FbsStep>& steps, std::vector<int2>& events,
                             int& depth_out, bool& stack_ok) {
    events.clear();
    depth_out = 0;
    stack_ok = true;
    // depth-first schedule for fbs_dfs_kernel: parent step of every step, children visited
    // latest-in-topo-order first, {step, 2*level + ascend} events.
    {
        std::vector<int> row_step((size_t)d->n_vrows, -2);
        for (int p = 0; p < 3; ++p) row_step[d->root_vrow[p]] = -1;
        for (int i = 0; i < d->n_steps; ++i)
            for (int p = 0; p < 3; ++p)
                if (steps[i].bus_vrow[p] >= 0) row_step[steps[i].bus_vrow[p]] = i;
        std::vector<std::vector<int>> kids((size_t)d->n_steps + 1);   // index 0 = root
        bool ok = true;
        for (int i = 0; i < d->n_steps && ok; ++i) {
            int par = -2;
            for (int p = 0; p < 3; ++p)
                if (steps[i].par_vrow[p] >= 0) par = row_step[steps[i].par_vrow[p]];
            if (par < -1 || par >= i) ok = false;                     // parents come first in topo order
            else kids[par + 1].push_back(i);
            if ((steps[i].kind & 0xff) == GP_EDGE_VR)                 // fbs_dfs_kernel: a regulator must cover every bus phase
                for (int p = 0; p < 3; ++p)
                    if (steps[i].bus_vrow[p] >= 0 && steps[i].gamma[p] == 0.0) stack_ok = false;
        }
        if (ok) {
            // Flags: what the parent must keep across this child's subtree.
            //  EV_VF   : the parent's forward voltage (another child still descends from it, or the
            //            first child does not rewrite every phase of the parent);
            //  EV_VBIA : the parent's rewritten voltage and current sum (not the first child).
            // A single child carrying all phases of its parent (a chain) saves nothing.
            auto covers = [&](int c) {        // child c rewrites every existing phase of its parent?
                for (int p = 0; p < 3; ++p)
                    if (steps[c].par_vrow[p] >= 0 && steps[c].bus_vrow[p] < 0) return false;
                if ((steps[c].kind & 0xff) == GP_EDGE_VR) return false;   // conservative
                return true;
            };
            struct Frame { int node, next; };
            std::vector<Frame> stack;
            stack.push_back({-1, 0});
            int depth = 0;
            for (auto& k : kids) std::sort(k.begin(), k.end(), std::greater<int>());
            while (!stack.empty()) {
                Frame& f = stack.back();
                auto& ks = kids[f.node + 1];
                const int m = (int)ks.size();
                if (f.next < m) {
                    const int j = f.next++;
                    const int c = ks[j];
                    const int level = (int)stack.size();
                    depth = level > depth ? level : depth;
                    const bool need_vf = (m > 1) || !covers(c);
                    int fl = 0;
                    if (j == 0 && need_vf) fl |= EV_VF;
                    if (j > 0) fl |= EV_VBIA;
                    events.push_back(make_int2(c, (level << 4) | fl));
                    stack.push_back({c, 0});
                } else {
                    if (f.node >= 0) {
                        // which child of its parent was this?
                        const Frame& pf = stack[stack.size() - 2];
                        const auto& pks = kids[pf.node + 1];
                        const int pm = (int)pks.size(), j = pf.next - 1;
                        const bool need_vf = (pm > 1) || !covers(f.node);
                        int fl = EV_ASCEND;
                        if (j < pm - 1 || (j == 0 && need_vf)) fl |= EV_VF;
                        if (j > 0) fl |= EV_VBIA;
                        events.push_back(make_int2(f.node, (((int)stack.size() - 1) << 4) | fl));
                    }
                    stack.pop_back();
                }
            }
            depth_out = depth;
            if ((int)events.size() != 2 * d->n_steps) events.clear();
            if (depth > 64) stack_ok = false;
        }
    }
}

// Host-only view of what the kernels are given for a feeder (no device needed): the depth-first event list
// and the per-step `kind` words with their flag bits (8..15 fast mask, 19..21 dead parent stores).  Used by the
// CPU tests, which run the kernel's loop over exactly these tables in NumPy.
extern "C" int gp_fbs_schedule(const GpFbsDesc* d, int32_t* events_out, int32_t* kind_out) {
    if (!d) return fail(GP_EINVAL, "gp_fbs_schedule: null descriptor");
    std::vector<FbsStep> steps;
    std::vector<int4> child4;
    std::vector<int> vrow_srow;
    std::vector<double> vrow_cap;
    FbsConst k;
    int rc = fbs_tables_from_desc(d, steps, child4, k, vrow_srow, vrow_cap);
    if (rc) return rc;
    std::vector<int2> events;
    int depth = 0;
    bool stack_ok = true;
    fbs_build_events(d, steps, events, depth, stack_ok);
    if (events_out)
        for (size_t e = 0; e < events.size(); ++e) {
            events_out[2 * e] = events[e].x;
            events_out[2 * e + 1] = events[e].y;
        }
    if (kind_out)
        for (size_t i = 0; i < steps.size(); ++i) kind_out[i] = steps[i].kind;
    return (int)events.size();
}


// Event table of fbs_walk_kernel from the step tables with shared I rows (steps_a: alias rows resolved, rows that
// stay zero dropped, bits 16-18 = rows this step does not store) - see the kernel for what each field means.
// Empty when the feeder cannot be walked this way (not a tree in topo order, a V row that no forward step writes).
static void fbs_build_walk(const GpFbsDesc* d, const std::vector<FbsStep>& steps, const std::vector<FbsStep>& steps_a,
                           int n_unfed, std::vector<WalkEv>& out) {
    out.clear();
    const int n = d->n_steps;
    if (n_unfed > 0 || n == 0) return;
    std::vector<int> row_step((size_t)d->n_vrows, -2);
    for (int p = 0; p < 3; ++p) row_step[d->root_vrow[p]] = -1;
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < 3; ++p)
            if (steps[i].bus_vrow[p] >= 0) row_step[steps[i].bus_vrow[p]] = i;
    std::vector<int> par((size_t)n, -2);
    std::vector<std::vector<int>> kids((size_t)n + 1);       // index 0 = root; children latest-in-topo-order first
    for (int i = 0; i < n; ++i) {
        for (int p = 0; p < 3; ++p)
            if (steps[i].par_vrow[p] >= 0) par[i] = row_step[steps[i].par_vrow[p]];
        if (par[i] < -1 || par[i] >= i) return;
        kids[par[i] + 1].push_back(i);
    }
    for (auto& ks : kids) std::sort(ks.begin(), ks.end(), std::greater<int>());
    // V rows of a bus (step or root) by phase
    auto vrow_of = [&](int node, int p) { return node < 0 ? d->root_vrow[p] : steps[node].bus_vrow[p]; };
    struct Frame { int node, next; };
    std::vector<Frame> stack;
    stack.push_back({-1, 0});
    auto base_event = [&](int c, int type) {
        WalkEv ev;
        memset(&ev, 0, sizeof(ev));
        const FbsStep& st = steps[c];
        const FbsStep& sa = steps_a[c];
        const bool vr = (st.kind & 0xff) == GP_EDGE_VR;
        ev.step = c;
        int mb = 0, me = 0, ms = 0, mw = 0, mc = 0, ml = 0;
        for (int p = 0; p < 3; ++p) {
            ev.e_row[p] = vr ? -1 : sa.edge_irow[p];
            ev.ld_i[p] = ev.ld_v[p] = ev.st_ca[p] = ev.ld_s[p] = -1;
            if (st.bus_vrow[p] >= 0) mb |= 1 << p;
            if (ev.e_row[p] >= 0) {
                me |= 1 << p;
                if (!((sa.kind >> (16 + p)) & 1)) ms |= 1 << p;
                // every line-kind child adds its edge current to the sum of its parent (:274-284); a parent
                // without that edge phase discards it
                if (par[c] >= 0 && (steps[par[c]].kind & 0xff) != GP_EDGE_VR && steps_a[par[c]].edge_irow[p] >= 0) mc |= 1 << p;
            }
            const bool stores = st.par_vrow[p] >= 0 && (vr ? st.gamma[p] != 0.0 : st.bus_vrow[p] >= 0);
            if (stores && !((st.kind >> (19 + p)) & 1)) mw |= 1 << p;
            if (st.bus_vrow[p] >= 0 && (st.load_srow[p] >= 0 || st.cap[p] != 0.0)) ml |= 1 << p;
            if (type != WK_DESC && !vr) ev.ld_s[p] = st.bus_vrow[p] >= 0 ? st.load_srow[p] : -1;
            if (type != WK_ASC) ev.ld_i[p] = ev.e_row[p];       // last iteration's edge current
        }
        int uni = 0;    // compile-time mask: bus phases = edge phases, all on the parent, a Line edge
        if (!vr && mb != 0 && mb == me && (mb == 7 || mb == 1 || mb == 2 || mb == 4)) {
            bool on_parent = true;
            for (int p = 0; p < 3; ++p)
                if (((mb >> p) & 1) && st.par_vrow[p] < 0) on_parent = false;
            if (on_parent) uni = mb;
        }
        ev.flags = type | (vr ? 4 : 0) | (mb << 3) | (ms << 6) | (mw << 9) | (mc << 12) | (ml << 15) | (me << 18) | (uni << 21);
        return ev;
    };
    // contributed[node + 1][p]: a child visited so far has added to phase p of the bus's current sum
    std::vector<int> contributed((size_t)n + 1, 0);
    while (!stack.empty()) {
        Frame& f = stack.back();
        auto& ks = kids[f.node + 1];
        if (f.next < (int)ks.size()) {
            const int c = ks[f.next++];
            const bool leaf = kids[c + 1].empty();
            WalkEv ev = base_event(c, leaf ? WK_LEAF : WK_DESC);
            const int k = f.node;
            if (leaf) {
                contributed[k + 1] |= (ev.flags >> 12) & 7;
                out.push_back(ev);
            } else {
                // park the parent's partial sum: owner rows of its edge that already hold contributions
                if (k >= 0)
                    for (int p = 0; p < 3; ++p)
                        if (((contributed[k + 1] >> p) & 1) && steps_a[k].edge_irow[p] >= 0 &&
                            !((steps_a[k].kind >> (16 + p)) & 1))
                            ev.st_ca[p] = steps_a[k].edge_irow[p];
                out.push_back(ev);
                stack.push_back({c, 0});
            }
        } else {
            const int c = f.node;
            stack.pop_back();
            if (c < 0) break;
            const int k = stack.back().node;
            WalkEv ev = base_event(c, WK_ASC);
            const int mw = (ev.flags >> 9) & 7;
            for (int p = 0; p < 3; ++p) {
                ev.ld_i[p] = -1;
                if (k >= 0 && ((contributed[k + 1] >> p) & 1)) ev.ld_i[p] = steps_a[k].edge_irow[p];   // parked sum / shared row
                const int vr = vrow_of(k, p);
                ev.ld_v[p] = (vr >= 0 && !((mw >> p) & 1)) ? vr : -1;
            }
            contributed[k + 1] |= (ev.flags >> 12) & 7;
            out.push_back(ev);
        }
    }
    if ((int)out.size() < n || (int)out.size() > 2 * n) out.clear();
}

// Host only: the event table of the tree-walk kernel for a descriptor, 20 int32 per event (struct WalkEv in
// gp_fbs.cu: step, flags, e_row[3], ld_i[3], ld_v[3], st_ca[3], ld_s[3], pad[3]) together with the step tables the
// kernel addresses (shared I rows resolved): edge_rows_out[n_steps][3].  Returns the number of events (0: the
// feeder is not walked) or a negative GP_E*.  The CPU tests run the kernel's loop over this table in NumPy.
static void fbs_shared_tables(const std::vector<FbsStep>& steps, const std::vector<int4>& child4, int n_irows,
                              std::vector<FbsStep>& steps_a, std::vector<int4>& child_a, std::vector<int2>& alias);
static void fbs_unfed_rows(const GpFbsDesc* d, const std::vector<FbsStep>& steps, std::vector<int>& unfed);

extern "C" int gp_fbs_walk_events(const GpFbsDesc* d, int32_t* events_out, int64_t cap_events) {
    if (!d) return fail(GP_EINVAL, "gp_fbs_walk_events: null descriptor");
    std::vector<FbsStep> steps, steps_a;
    std::vector<int4> child4, child_a;
    std::vector<int2> alias;
    std::vector<int> vrow_srow, unfed;
    std::vector<double> vrow_cap;
    FbsConst k;
    int rc = fbs_tables_from_desc(d, steps, child4, k, vrow_srow, vrow_cap);
    if (rc) return rc;
    fbs_shared_tables(steps, child4, d->n_irows, steps_a, child_a, alias);
    fbs_unfed_rows(d, steps, unfed);
    std::vector<WalkEv> ev;
    fbs_build_walk(d, steps, steps_a, (int)unfed.size(), ev);
    if (events_out) {
        if ((int64_t)ev.size() > cap_events) return fail(GP_EINVAL, "gp_fbs_walk_events: buffer too small");
        memcpy(events_out, ev.data(), ev.size() * sizeof(WalkEv));
    }
    return (int)ev.size();
}

// Everything an FbsImpl owns (device tables, modules, staging, streams); also the error path of gp_fbs_create.
static void fbs_impl_free(FbsImpl* im) {
    if (!im) return;
    cudaFree(im->d_steps);
    cudaFree(im->d_child);
    cudaFree(im->d_steps_a);
    cudaFree(im->d_child_a);
    cudaFree(im->d_alias);
    cudaFree(im->d_lines);
    cudaFree(im->d_vrow_srow);
    cudaFree(im->d_vrow_cap);
    cudaFree(im->d_events);
    cudaFree(im->d_evrows);
    cudaFree(im->d_sched);
    cudaFree(im->d_unfed);
    cudaFree(im->d_walk);
    cudaFree(im->d_zc);
    if (im->jit_module) driver().ModuleUnload(im->jit_module);
    if (im->jit_module_host) driver().ModuleUnload(im->jit_module_host);
    if (im->jit_module_ex) driver().ModuleUnload(im->jit_module_ex);
    for (int i = 0; i < FbsImpl::NSTREAM; ++i) {
        if (im->d_stage[i]) cudaFree(im->d_stage[i]);
        if (im->streams[i]) cudaStreamDestroy(im->streams[i]);
    }
    delete im;
}

extern "C" int gp_feeder_destroy(GpFeeder* f) {
    if (!f) return GP_OK;
    cudaSetDevice(f->device);
    if (f->kind == 1) {
        fbs_impl_free((FbsImpl*)f->impl);
    } else if (f->kind == 2) {
        gp_nr3_destroy_impl(f->impl);
    }
    delete f;
    return GP_OK;
}


// The streaming kernels' tables with shared I rows (fbs_jit_islots has the rule): an I row of a bus-phase without
// load and capacitor that is fed by exactly one child row IS that row's value (0 + I_child), so its forward
// read goes to the child's row and its backward store is dropped (bits 16..18 of `kind`); one that nothing feeds
// stays zero and leaves the tables altogether (the step then takes the generic body, which skips absent rows).
// Per sweep that is one write and one read less per such row; `alias` lists {row, source row or -1} for the one
// copy-out after the last sweep.
static void fbs_shared_tables(const std::vector<FbsStep>& steps, const std::vector<int4>& child4, int n_irows,
                              std::vector<FbsStep>& steps_a, std::vector<int4>& child_a, std::vector<int2>& alias) {
    std::vector<int> slot;
    std::vector<char> owner;
    fbs_jit_islots(steps, child4, n_irows, slot, owner);
    std::vector<int> owner_row(slot.size() + 1, -1);
    for (size_t r = 0; r < slot.size(); ++r)
        if (owner[r] && slot[r] >= 0) owner_row[slot[r]] = (int)r;
    auto remap = [&](int r) { return (r < 0 || r >= n_irows) ? r : (slot[r] < 0 ? -1 : owner_row[slot[r]]); };
    steps_a = steps;
    child_a = child4;
    alias.clear();
    for (FbsStep& st : steps_a)
        for (int p = 0; p < 3; ++p) {
            const int r = st.edge_irow[p];
            if (r < 0 || r >= n_irows || owner[r]) continue;
            alias.push_back(make_int2(r, remap(r)));
            st.edge_irow[p] = remap(r);
            if (st.edge_irow[p] < 0) st.kind &= ~0xff00;        // a dropped row: no mask-specialised body
            else st.kind |= 1 << (16 + p);
        }
    for (int4& c : child_a) c = make_int4(remap(c.x), remap(c.y), remap(c.z), c.w);
}

// V rows that no forward step writes (non-regulated phases below a regulator): they must hold the flat start's zero
static void fbs_unfed_rows(const GpFbsDesc* d, const std::vector<FbsStep>& steps, std::vector<int>& unfed) {
    std::vector<char> fed((size_t)d->n_vrows, 0);
    for (int p = 0; p < 3; ++p) fed[d->root_vrow[p]] = 1;
    for (const FbsStep& st : steps)
        for (int p = 0; p < 3; ++p)
            if (st.bus_vrow[p] >= 0 && ((st.kind & 0xff) != GP_EDGE_VR || st.gamma[p] != 0.0)) fed[st.bus_vrow[p]] = 1;
    unfed.clear();
    for (int r = 0; r < d->n_vrows; ++r)
        if (!fed[r]) unfed.push_back(r);
}

extern "C" int gp_fbs_create(const GpFbsDesc* d, int device, GpFeeder** out) {
    if (!d || !out) return fail(GP_EINVAL, "gp_fbs_create: null argument");
    std::vector<FbsStep> steps;
    std::vector<int4> child4;
    std::vector<int> vrow_srow;
    std::vector<double> vrow_cap;
    FbsConst k0;
    int rc0 = fbs_tables_from_desc(d, steps, child4, k0, vrow_srow, vrow_cap);
    if (rc0) return rc0;
    GP_CUDA(cudaSetDevice(device));
    FbsImpl* im = new (std::nothrow) FbsImpl();
    if (!im) return fail(GP_ENOMEM, "gp_fbs_create: out of host memory");
    im->n_steps = d->n_steps;
    im->n_vrows = d->n_vrows;
    im->n_irows = d->n_irows;
    im->n_srows = d->n_srows;
    im->n_child = d->n_child;
    im->n_lines = d->n_lines;
    for (const FbsStep& st : steps)
        for (int p = 0; p < 3; ++p) im->n_grows += st.gam_row[p] >= 0;
    im->k = k0;
    im->h_steps = steps;
    im->h_child = child4;
    std::vector<int> lines((size_t)d->n_lines * 9);
    for (int l = 0; l < d->n_lines; ++l)
        for (int p = 0; p < 3; ++p) {
            lines[9 * l + p] = d->line_irow[3 * l + p];
            lines[9 * l + 3 + p] = d->line_tx_vrow[3 * l + p];
            lines[9 * l + 6 + p] = d->line_rx_vrow[3 * l + p];
        }
    im->h_lines = lines;
    std::vector<int2> events;
    fbs_build_events(d, steps, events, im->depth, im->dfs_stack_ok);
    im->n_events = (int)events.size();
    std::vector<int4> evrows(events.size());
    for (size_t e = 0; e < events.size(); ++e) {
        const FbsStep& st = steps[events[e].x];
        const bool asc = (events[e].y & EV_ASCEND) != 0;
        const int* r = asc ? st.load_srow : st.edge_irow;
        if ((st.kind & 0xff) == GP_EDGE_VR) evrows[e] = make_int4(-1, -1, -1, asc ? 0 : 1);
        else evrows[e] = make_int4(r[0], r[1], r[2], asc ? 0 : 1);
    }
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes ? bytes : 16);
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t e = up((void**)&im->d_steps, steps.data(), steps.size() * sizeof(FbsStep));
    if (e == cudaSuccess) e = up((void**)&im->d_events, events.data(), events.size() * sizeof(int2));
    if (e == cudaSuccess) e = up((void**)&im->d_evrows, evrows.data(), evrows.size() * sizeof(int4));
    if (e == cudaSuccess) e = up((void**)&im->d_child, child4.data(), child4.size() * sizeof(int4));
    if (e == cudaSuccess) e = up((void**)&im->d_lines, lines.data(), lines.size() * sizeof(int));
    std::vector<FbsStep> steps_a;
    std::vector<int4> child_a;
    {
        std::vector<int2> alias;
        fbs_shared_tables(steps, child4, d->n_irows, steps_a, child_a, alias);
        im->n_alias = (int)alias.size();
        if (e == cudaSuccess) e = up((void**)&im->d_steps_a, steps_a.data(), steps_a.size() * sizeof(FbsStep));
        if (e == cudaSuccess) e = up((void**)&im->d_child_a, child_a.data(), child_a.size() * sizeof(int4));
        if (e == cudaSuccess) e = up((void**)&im->d_alias, alias.data(), alias.size() * sizeof(int2));
    }
    if (e == cudaSuccess) e = up((void**)&im->d_vrow_srow, vrow_srow.data(), vrow_srow.size() * sizeof(int));
    if (e == cudaSuccess) e = up((void**)&im->d_vrow_cap, vrow_cap.data(), vrow_cap.size() * sizeof(double));
    {
        std::vector<int> unfed;
        fbs_unfed_rows(d, steps, unfed);
        im->n_unfed = (int)unfed.size();
        if (e == cudaSuccess) e = up((void**)&im->d_unfed, unfed.data(), unfed.size() * sizeof(int));
        std::vector<WalkEv> walk;
        fbs_build_walk(d, steps, steps_a, im->n_unfed, walk);
        std::vector<WalkRec> recs(walk.size());     // the 64-byte records the kernel reads
        for (size_t i = 0; i < walk.size(); ++i) {
            const WalkEv& w = walk[i];
            const bool asc = (w.flags & 3) == WK_ASC;
            const int* c = asc ? w.ld_i : w.st_ca;
            int fl = w.flags & 0x00ffffff;
            if (!asc && (w.st_ca[0] >= 0 || w.st_ca[1] >= 0 || w.st_ca[2] >= 0)) fl |= 1 << 24;
            recs[i].w[0] = make_int4(w.step, fl, w.e_row[0], w.e_row[1]);
            recs[i].w[1] = make_int4(w.e_row[2], w.ld_s[0], w.ld_s[1], w.ld_s[2]);
            recs[i].w[2] = make_int4(c[0], c[1], c[2], asc ? w.ld_v[0] : -1);
            recs[i].w[3] = make_int4(asc ? w.ld_v[1] : -1, asc ? w.ld_v[2] : -1, 0, 0);
        }
        im->n_walk = (int)recs.size();
        if (e == cudaSuccess) e = up((void**)&im->d_walk, recs.data(), recs.size() * sizeof(WalkRec));
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&im->d_sched, (FbsImpl::NSCHED + 1) * 16);
    if (e == cudaSuccess) e = cudaMemset(im->d_sched, 0, (FbsImpl::NSCHED + 1) * 16);
    int smem_optin = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&im->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) {
        fbs_impl_free(im);
        return fail(GP_ECUDA, "gp_fbs_create: %s", cudaGetErrorString(e));
    }
    {   // warps per block of the on-chip kernel: one 32-scenario slice of V and I rows per warp
        const size_t per_warp = (size_t)(d->n_vrows + d->n_irows) * 512;
        const int fit = (int)((size_t)smem_optin / per_warp);
        static const int sizes[] = {16, 12, 8, 6, 4, 3, 2, 1};
        for (int w : sizes)
            if (w <= fit) {
                im->onchip_warps = w;
                break;
            }
    }
    {   // plan of the specialised kernel: I rows in shared memory next to V, or in global memory (the I
        // array) when that lets more warps run per SM; at most 8 warps (about 240 registers per thread).
        // In shared memory the I rows take one slot per DISTINCT value (fbs_jit_islots).
        {
            std::vector<int> slot;
            std::vector<char> owner;
            im->n_islots = fbs_jit_islots(im->h_steps, im->h_child, d->n_irows, slot, owner);
        }
        static const int sizes[] = {8, 7, 6, 5, 4, 3, 2, 1};
        auto pick = [&](size_t per_warp) {
            const int fit = (int)((size_t)smem_optin / per_warp);
            for (int w : sizes)
                if (w <= fit) return w;
            return 0;
        };
        const int w_is = pick((size_t)(d->n_vrows + im->n_islots) * 512), w_ig = pick((size_t)d->n_vrows * 512);
        // measured on B200: 13-bus 6 -> 8 warps: no gain (the register-passed currents spill);
        // 34-bus 2 -> 4 warps: 1.21x; 37-bus 1 -> 3 warps: 1.49x  =>  only when it at least doubles the warps
        im->jit_plan_ig = g_jit_i_global == 1 || (g_jit_i_global < 0 && w_ig >= 2 * w_is && w_ig > 0);
        im->jit_plan_warps = im->jit_plan_ig ? w_ig : w_is;
        // depth-first build: the V rows leave shared memory altogether (I slots + a small stack stay).  Automatic when
        // it at least doubles the warps of the two-sweep build (34-bus 3 -> 6, 37-bus 2 -> 4; 13-bus 7 -> 8: no)
        im->jit_alt_warps = im->jit_plan_warps;     // the two-sweep plan: what the one-launch host path runs when the
        im->jit_alt_ig = im->jit_plan_ig;           // device path takes the depth-first build
        im->n_dfs_stack = fbs_jit_dfs_stack(im->h_steps, im->h_child, im->k, d->n_vrows, d->n_irows);
        if (im->n_dfs_stack >= 0 && g_jit_dfs != 0) {
            const int w_dfs = pick((size_t)(im->n_islots + im->n_dfs_stack) * 512);
            if (w_dfs > 0 && (g_jit_dfs == 1 || w_dfs >= 2 * im->jit_plan_warps)) {
                im->jit_plan_dfs = true;
                im->jit_plan_ig = false;
                im->jit_plan_warps = w_dfs;
            }
        }
    }
    GpFeeder* f = new (std::nothrow) GpFeeder();
    if (!f) {
        fbs_impl_free(im);
        return fail(GP_ENOMEM, "gp_fbs_create: out of host memory");
    }
    f->device = device;
    f->kind = 1;
    f->impl = im;
    *out = f;
    return GP_OK;
}

static int fbs_check(GpFeeder* f, int64_t S, int64_t ld, const char* who) {
    if (!f || f->kind != 1) return fail(GP_EINVAL, "%s: not an FBS feeder handle", who);
    if (S < 0 || ld < S) return fail(GP_EINVAL, "%s: need 0 <= S <= ld (S=%lld ld=%lld)", who, (long long)S, (long long)ld);
    return GP_OK;
}

// Scheduler words {next tile, warps done} of a launch of a persistent FBS kernel on stream `st`.  The words
// re-arm themselves when the last warp of a launch leaves, which is only correct for ONE kernel in flight
// per pair: every stream gets its own pair (launches on one stream are ordered).  A handle that has seen
// more than NSCHED streams shares the overflow pair, which is zeroed on the launch stream after waiting
// for that pair's previous launch.
static int sched_words(FbsImpl* im, cudaStream_t st, unsigned** out) {
    for (int i = 0; i < im->n_sched_streams; ++i)
        if (im->sched_stream[i] == st) {
            *out = im->d_sched + 4 * i;
            return GP_OK;
        }
    if (im->n_sched_streams < FbsImpl::NSCHED) {
        im->sched_stream[im->n_sched_streams] = st;
        *out = im->d_sched + 4 * im->n_sched_streams++;
        return GP_OK;
    }
    GP_CUDA(cudaDeviceSynchronize());
    GP_CUDA(cudaMemsetAsync(im->d_sched + 4 * FbsImpl::NSCHED, 0, 16, st));
    *out = im->d_sched + 4 * FbsImpl::NSCHED;
    return GP_OK;
}

// Launch of the feeder-specialised kernel.  `stage` != NULL: spu is mapped host memory and is staged
// through these device rows by the kernel itself; V, I, Vmag, loss, iters, status, diff may then be
// mapped host memory too (the kernel's final stores go straight over PCIe).
static int fbs_jit_launch(FbsImpl* im, int64_t S, int64_t ld, const double* spu, double* V, double* I,
                          int32_t* iters, int32_t* status, double* diff, double tol, int32_t maxiter,
                          cudaStream_t st, double* Vmag, double* loss, void* stage, double* Iout,
                          const PeerDst* peers = nullptr, void* kernel = nullptr, int kernel_warps = 0,
                          const char* gam = nullptr, int warm = 0, int kernel_rows = 0) {
    const long long n_tiles = (S + 31) / 32;
    const int nw = (kernel && kernel_warps > 0) ? kernel_warps : im->jit_warps;   // a build's own block size
    long long blocks = (n_tiles + nw - 1) / nw;
    if (blocks > im->n_sm) blocks = im->n_sm;
    const unsigned smem = (unsigned)nw * (unsigned)((kernel && kernel_rows > 0) ? kernel_rows : im->jit_rows) * 512u;
    long long S_arg = S;
    double tol_arg = tol;
    int maxiter_arg = maxiter;
    unsigned ldb_arg = (unsigned)(ld * 16);
    PeerDst pd;
    memset(&pd, 0, sizeof(pd));
    if (peers) pd = *peers;
    unsigned* sched = nullptr;
    int rcs = sched_words(im, st, &sched);
    if (rcs) return rcs;
    void* args[] = {(void*)&spu, (void*)&V, (void*)&I, (void*)&iters, (void*)&status, (void*)&diff,
                    (void*)&S_arg, (void*)&ldb_arg, (void*)&tol_arg, (void*)&maxiter_arg, (void*)&sched,
                    (void*)&Vmag, (void*)&loss, (void*)&stage, (void*)&Iout, (void*)&pd,
                    (void*)&gam, (void*)&warm};        // the last two are parameters of the EX build only
    const int cr = driver().LaunchKernel(kernel ? kernel : im->jit_kernel, (unsigned)blocks, 1, 1, (unsigned)nw * 32, 1, 1, smem,
                                         (void*)st, args, nullptr);
    if (cr != 0) {
        cudaMemsetAsync(sched, 0, 16, st);      // a launch that never ran must not leave the words armed
        return fail(GP_ECUDA, "gp_fbs_solve_batch: cuLaunchKernel failed (%d)", cr);
    }
    return GP_OK;
}

// GpSolveExtras.peer_* -> the by-value kernel argument
static int peers_from_extras(const GpSolveExtras* ex, PeerDst& pd) {
    memset(&pd, 0, sizeof(pd));
    if (!ex || ex->n_peers == 0) return GP_OK;
    if (ex->n_peers < 0 || ex->n_peers > GP_MAX_PEERS || !ex->peer_vmag || !ex->peer_loss)
        return fail(GP_EINVAL, "GpSolveExtras: n_peers must be 0..%d with both pointer arrays set", GP_MAX_PEERS);
    for (int q = 0; q < ex->n_peers; ++q) {
        if (!ex->peer_vmag[q] || !ex->peer_loss[q]) return fail(GP_EINVAL, "GpSolveExtras: null peer destination %d", q);
        pd.vmag[q] = ex->peer_vmag[q];
        pd.loss[q] = (double2*)ex->peer_loss[q];
    }
    pd.n = ex->n_peers;
    return GP_OK;
}

// GpSolveExtras.vv_* -> the droop fields of the FbsConst copy a launch passes (the ramps' slope and intercept with
// VoltVARController.get_Q's own operations, volt_var_controller.py:25-33)
static int droop_from_extras(const GpSolveExtras* ex, FbsConst& k, const int** vv) {
    *vv = nullptr;
    for (int i = 0; i < 4; ++i) k.vv_v[i] = 0.0;
    k.vv_lo[0] = k.vv_lo[1] = k.vv_hi[0] = k.vv_hi[1] = k.vv_q[0] = k.vv_q[1] = 0.0;
    if (!ex || !ex->vv_ctl_dev) return GP_OK;
    const double V1 = ex->vv_points[0], V2 = ex->vv_points[1], V3 = ex->vv_points[2], V4 = ex->vv_points[3];
    if (!(V1 < V2 && V2 <= V3 && V3 < V4))
        return fail(GP_EINVAL, "GpSolveExtras: volt-var breakpoints need V1 < V2 <= V3 < V4");
    const double minQ = ex->vv_q[0], maxQ = ex->vv_q[1];
    k.vv_v[0] = V1; k.vv_v[1] = V2; k.vv_v[2] = V3; k.vv_v[3] = V4;
    k.vv_lo[0] = (-maxQ) / (V2 - V1);
    k.vv_lo[1] = -(k.vv_lo[0] * V2);
    k.vv_hi[0] = (minQ) / (V4 - V3);
    k.vv_hi[1] = -(k.vv_hi[0] * V3);
    k.vv_q[0] = minQ;
    k.vv_q[1] = maxQ;
    *vv = ex->vv_ctl_dev;
    return GP_OK;
}

// Launch the solve kernel; *fused is set when the kernel also produced Vmag / loss.
static int fbs_solve_launch(GpFeeder* f, int64_t S, int64_t ld, const double* spu, double* V, double* I,
                            int32_t* iters, int32_t* status, double* diff, double tol, int32_t maxiter,
                            void* stream, double* Vmag, double* loss, bool* fused,
                            const GpSolveExtras* ex = nullptr) {
    *fused = false;
    int rc = fbs_check(f, S, ld, "gp_fbs_solve_batch");
    if (rc) return rc;
    if (S == 0) return GP_OK;
    FbsImpl* im = (FbsImpl*)f->impl;
    const char* gam = ex ? (const char*)ex->gamma_dev : nullptr;
    const int warm = ex ? ex->warm_start : 0;
    const char* wpu = ex ? (const char*)ex->wpu_dev : nullptr;
    if (warm != 0 && warm != 1) return fail(GP_EINVAL, "gp_fbs_solve_batch_ex: warm_start must be 0 or 1");
    if (gam && im->n_grows == 0) gam = nullptr;     // no regulators: nothing to read
    FbsConst kx = im->k;
    const int* vv = nullptr;
    rc = droop_from_extras(ex, kx, &vv);
    if (rc) return rc;
    const bool extras = gam || warm || wpu || vv;
    PeerDst pd;
    rc = peers_from_extras(ex, pd);
    if (rc) return rc;
    if (!V || !I || !iters || !status || !diff || (!spu && im->n_srows > 0))
        return fail(GP_EINVAL, "gp_fbs_solve_batch: null buffer");
    if (ld * 16 >= ((int64_t)1 << 32))
        return fail(GP_EINVAL, "gp_fbs_solve_batch: row pitch ld must be below 2^28 scenarios; split the batch");
    GP_CUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int BLOCK = 128;
    const unsigned gx = (unsigned)((S + BLOCK - 1) / BLOCK);
    const bool dfs = im->n_events > 0 && im->dfs_stack_ok && g_fbs_kernel == 1;
    const unsigned ldb = (unsigned)(ld * 16);
    // automatic choice: the feeder-specialised on-chip kernel when gp_fbs_specialise has built it,
    // else the two-sweep kernel that streams the state through HBM.  (The table-driven on-chip
    // kernel is slower than streaming on B200 - interpretation overhead with only a handful of
    // warps per SM - and runs only on request.)
    const bool onchip = g_fbs_kernel == 4;
    // (the specialised kernel always runs at least one sweep: maxiter < 1 goes to the table-driven ones)
    // (a handle whose plain kernel is the depth-first build takes the two-sweep plan saved at create for this build)
    const bool ex_fits = im->jit_kernel && (im->jit_dfs ? (im->jit_alt_warps >= 1 && !im->jit_alt_ig) : !im->jit_ig);
    const int ex_warps = im->jit_dfs ? im->jit_alt_warps : im->jit_warps;
    if (extras && !wpu && !vv && ex_fits && maxiter >= 1 && (g_fbs_kernel == 0 || g_fbs_kernel == 5)) {
        // per-scenario ratios / warm start on a specialised feeder: the EX build of the specialised kernel
        // (built on first use; a failed build leaves the streaming kernel's EX instance below)
        if (!im->jit_ex_tried) {
            im->jit_ex_tried = true;
            if (fbs_jit_build(f, ex_warps, false, false, &im->jit_module_ex, &im->jit_kernel_ex, true) != GP_OK)
                im->jit_module_ex = im->jit_kernel_ex = nullptr;
        }
    }
    if (extras && !wpu && !vv && im->jit_kernel_ex && ex_fits && maxiter >= 1 &&
        (g_fbs_kernel == 0 || g_fbs_kernel == 5)) {
        rc = fbs_jit_launch(im, S, ld, spu, V, I, iters, status, diff, tol, maxiter, st, Vmag, loss, nullptr, nullptr,
                            &pd, im->jit_kernel_ex, ex_warps, gam, warm, im->n_vrows + im->n_islots);
        if (rc) return rc;
        *fused = true;
    } else if (extras) {
        // per-scenario ratios / warm start / wpu / droop: the streaming kernel's EX instance (the table-driven
        // on-chip kernel bakes the ratios into its tables and always starts flat)
        const int pf = g_fbs_prefetch >= 0 ? g_fbs_prefetch : 0;
        if (g_fbs_order && im->n_events > 0)
            fbs_solve_kernel<BLOCK, true, true><<<gx, BLOCK, 0, st>>>(
                im->d_steps, im->d_child, kx, im->n_steps, im->n_vrows, im->n_irows, S, ldb, (const char*)spu, (char*)V,
                (char*)I, iters, status, diff, tol, maxiter, pf, im->d_unfed, im->n_unfed, gam, warm, wpu, nullptr, 0,
                im->d_events, im->n_events, vv);
        else
            fbs_solve_kernel<BLOCK, true, false><<<gx, BLOCK, 0, st>>>(
                im->d_steps, im->d_child, kx, im->n_steps, im->n_vrows, im->n_irows, S, ldb, (const char*)spu, (char*)V,
                (char*)I, iters, status, diff, tol, maxiter, pf, im->d_unfed, im->n_unfed, gam, warm, wpu, nullptr, 0,
                nullptr, 0, vv);
    } else if ((g_fbs_kernel == 5 || (g_fbs_kernel == 0 && im->jit_kernel)) && (maxiter >= 1 || !im->jit_kernel)) {
        if (!im->jit_kernel)
            return fail(GP_EINVAL, "gp_fbs_solve_batch: fbs_kernel = 5 needs a successful gp_fbs_specialise");
        rc = fbs_jit_launch(im, S, ld, spu, V, I, iters, status, diff, tol, maxiter, st, Vmag, loss, nullptr, nullptr,
                            &pd);
        if (rc) return rc;
        *fused = true;
    } else if (onchip) {
        if (im->onchip_warps < 1)
            return fail(GP_EINVAL, "gp_fbs_solve_batch: the on-chip kernel needs %d bytes of shared memory per warp",
                        (im->n_vrows + im->n_irows) * 512);
        const long long n_tiles = (S + 31) / 32;
        const int nw = im->onchip_warps;
        const size_t smem = (size_t)nw * (size_t)(im->n_vrows + im->n_irows) * 512;
        unsigned* sched = nullptr;
        rc = sched_words(im, st, &sched);
        if (rc) return rc;
        long long blocks = (n_tiles + nw - 1) / nw;
        if (blocks > im->n_sm) blocks = im->n_sm;
#define GP_ONCHIP(NW)                                                                                           \
    case NW: {                                                                                                  \
        static bool attr_set = false;   /* per template instance; the limit is per function, any device */       \
        if (!attr_set) {                                                                                        \
            GP_CUDA(cudaFuncSetAttribute(fbs_onchip_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         232448));                                                              \
            attr_set = true;                                                                                    \
        }                                                                                                       \
        fbs_onchip_kernel<NW><<<(unsigned)blocks, NW * 32, smem, st>>>(                                         \
            im->d_steps, im->d_child, im->k, im->n_steps, im->n_vrows, im->n_irows, S, ldb, (const char*)spu,   \
            (char*)V, (char*)I, iters, status, diff, tol, maxiter, sched);                                      \
        break;                                                                                                  \
    }
        switch (nw) {
            GP_ONCHIP(16) GP_ONCHIP(12) GP_ONCHIP(8) GP_ONCHIP(6) GP_ONCHIP(4) GP_ONCHIP(3) GP_ONCHIP(2) GP_ONCHIP(1)
            default: return fail(GP_EINVAL, "gp_fbs_solve_batch: bad on-chip warp count %d", nw);
        }
#undef GP_ONCHIP
    } else if (g_fbs_kernel == 2) {
        const long long warps = (S + 9) / 10;
        const unsigned gp = (unsigned)((warps * 32 + BLOCK - 1) / BLOCK);
        fbs_phase_kernel<BLOCK><<<gp, BLOCK, 0, st>>>(im->d_steps, im->d_child, im->k, im->n_steps, im->n_vrows,
                                                      im->n_irows, S, ldb, (const char*)spu, (char*)V,
                                                      (char*)I, iters, status, diff, tol, maxiter);
    } else if (dfs) {
#define GP_DFS(D)                                                                                          \
    fbs_dfs_kernel<BLOCK, D><<<gx, BLOCK, 0, st>>>(im->d_steps, im->d_events, im->d_evrows, im->n_events, im->k, S, ldb, \
                                                   (const char*)spu, (char*)V, (char*)I, iters, status,    \
                                                   diff, tol, maxiter)
        if (im->depth <= 8) GP_DFS(8);
        else if (im->depth <= 16) GP_DFS(16);
        else if (im->depth <= 32) GP_DFS(32);
        else GP_DFS(64);
#undef GP_DFS
    } else if ((g_fbs_kernel == 6 || (g_fbs_kernel == 0 && g_fbs_auto_walk)) && im->n_walk > 0 && maxiter >= 1) {
        // tree walk with the state of the current bus in registers (always on the shared-row tables)
#define GP_WALK(MINB)                                                                                                   \
    fbs_walk_kernel<BLOCK, MINB><<<gx, BLOCK, 0, st>>>(im->d_steps_a, (const WalkRec*)im->d_walk, im->n_walk, im->k,     \
                                                       im->n_vrows, im->n_irows, S, ldb, (const char*)spu, (char*)V,    \
                                                       (char*)I, iters, status, diff, tol, maxiter, im->d_alias,        \
                                                       im->n_alias)
        switch (g_fbs_walk_blocks) {
            case 8: GP_WALK(8); break;
            case 6: GP_WALK(6); break;
            case 5: GP_WALK(5); break;
            default: GP_WALK(4); break;
        }
#undef GP_WALK
    } else {
        if (g_fbs_kernel == 6) return fail(GP_EINVAL, "gp_fbs_solve_batch: fbs_kernel = 6 needs a feeder the tree-walk kernel takes");
        const bool share = g_fbs_share_rows && im->n_alias > 0;
        const int pf = g_fbs_prefetch >= 0 ? g_fbs_prefetch : 0;       // see g_fbs_prefetch: every distance measured slower
        const FbsStep* tb = share ? im->d_steps_a : im->d_steps;
        const int4* tc = share ? im->d_child_a : im->d_child;
        if (g_fbs_order && im->n_events > 0)
            fbs_solve_kernel<BLOCK, false, true><<<gx, BLOCK, 0, st>>>(
                tb, tc, im->k, im->n_steps, im->n_vrows, im->n_irows, S, ldb, (const char*)spu, (char*)V, (char*)I, iters,
                status, diff, tol, maxiter, pf, im->d_unfed, im->n_unfed, nullptr, 0, nullptr,
                share ? im->d_alias : nullptr, share ? im->n_alias : 0, im->d_events, im->n_events, nullptr);
        else
            fbs_solve_kernel<BLOCK, false, false><<<gx, BLOCK, 0, st>>>(
                tb, tc, im->k, im->n_steps, im->n_vrows, im->n_irows, S, ldb, (const char*)spu, (char*)V, (char*)I, iters,
                status, diff, tol, maxiter, pf, im->d_unfed, im->n_unfed, nullptr, 0, nullptr,
                share ? im->d_alias : nullptr, share ? im->n_alias : 0, nullptr, 0, nullptr);
    }
    count_launch();
    GP_CUDA(cudaGetLastError());
    return GP_OK;
}

extern "C" int gp_fbs_solve_batch(GpFeeder* f, int64_t S, int64_t ld, const double* spu,
                                  double* V, double* I, int32_t* iters, int32_t* status,
                                  double* diff, double tol, int32_t maxiter, void* stream) {
    bool fused;
    return fbs_solve_launch(f, S, ld, spu, V, I, iters, status, diff, tol, maxiter, stream, nullptr, nullptr, &fused);
}

extern "C" int gp_fbs_post(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const double* V,
                           const double* I, double* sV, double* Vmag, double* Stx, double* Srx,
                           double* loss, void* stream);

extern "C" int gp_fbs_solve_post_batch(GpFeeder* f, int64_t S, int64_t ld, const double* spu, double* V, double* I,
                                       int32_t* iters, int32_t* status, double* diff, double* Vmag, double* loss,
                                       double tol, int32_t maxiter, void* stream) {
    bool fused;
    int rc = fbs_solve_launch(f, S, ld, spu, V, I, iters, status, diff, tol, maxiter, stream, Vmag, loss, &fused);
    if (rc || fused || S == 0 || (!Vmag && !loss)) return rc;
    return gp_fbs_post(f, S, ld, spu, V, I, nullptr, Vmag, nullptr, nullptr, loss, stream);
}

static int fbs_post_impl(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const double* V, const double* I,
                         double* sV, double* Vmag, double* Stx, double* Srx, double* loss, void* stream,
                         const PeerDst& pd, const double* wpu = nullptr, const GpSolveExtras* ex = nullptr);

extern "C" int gp_fbs_solve_batch_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const GpSolveExtras* ex,
                                     double* V, double* I, int32_t* iters, int32_t* status, double* diff,
                                     double* Vmag, double* loss, double tol, int32_t maxiter, void* stream) {
    bool fused;
    int rc = fbs_solve_launch(f, S, ld, spu, V, I, iters, status, diff, tol, maxiter, stream, Vmag, loss, &fused, ex);
    PeerDst pd;
    if (!rc) rc = peers_from_extras(ex, pd);
    if (rc || fused || S == 0 || (!Vmag && !loss && pd.n == 0)) return rc;
    return fbs_post_impl(f, S, ld, spu, V, I, nullptr, Vmag, nullptr, nullptr, loss, stream, pd);
}

extern "C" int gp_gamma_rows(GpFeeder* f) {
    if (!f || !f->impl) return fail(GP_EINVAL, "gp_gamma_rows: null handle");
    return f->kind == 1 ? ((FbsImpl*)f->impl)->n_grows : nr3_gamma_rows(f);
}

extern "C" int gp_fbs_post(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const double* V,
                           const double* I, double* sV, double* Vmag, double* Stx, double* Srx,
                           double* loss, void* stream) {
    PeerDst pd;
    memset(&pd, 0, sizeof(pd));
    return fbs_post_impl(f, S, ld, spu, V, I, sV, Vmag, Stx, Srx, loss, stream, pd);
}

extern "C" int gp_fbs_post_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const GpSolveExtras* ex,
                              const double* V, const double* I, double* sV, double* Vmag, double* Stx, double* Srx,
                              double* loss, void* stream) {
    PeerDst pd;
    memset(&pd, 0, sizeof(pd));
    return fbs_post_impl(f, S, ld, spu, V, I, sV, Vmag, Stx, Srx, loss, stream, pd, ex ? ex->wpu_dev : nullptr, ex);
}

static int fbs_post_impl(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const double* V, const double* I,
                         double* sV, double* Vmag, double* Stx, double* Srx, double* loss, void* stream,
                         const PeerDst& pd, const double* wpu, const GpSolveExtras* ex) {
    int rc = fbs_check(f, S, ld, "gp_fbs_post");
    if (rc) return rc;
    FbsConst kx = ((FbsImpl*)f->impl)->k;
    const int* vv = nullptr;
    rc = droop_from_extras(ex, kx, &vv);
    if (rc) return rc;
    if (S == 0) return GP_OK;
    FbsImpl* im = (FbsImpl*)f->impl;
    if (!V || !I) return fail(GP_EINVAL, "gp_fbs_post: null state buffer");
    if (sV && !spu && im->n_srows > 0) return fail(GP_EINVAL, "gp_fbs_post: sV needs spu");
    GP_CUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int BLOCK = 128;
    const unsigned gx = (unsigned)((S + BLOCK - 1) / BLOCK);
    if (sV || Vmag || pd.n) {
        dim3 g(gx, 8);
        fbs_post_rows_kernel<<<g, BLOCK, 0, st>>>(im->n_vrows, im->d_vrow_srow, im->d_vrow_cap, kx, S, ld,
                                                  (const double2*)spu, (const double2*)V, (double2*)sV, Vmag, pd, wpu, vv);
        count_launch();
    }
    if (Stx || Srx || loss || pd.n) {
        fbs_post_lines_kernel<<<gx, BLOCK, 0, st>>>(im->n_lines, im->d_lines, S, ld, (const double2*)V,
                                                    (const double2*)I, (double2*)Stx, (double2*)Srx,
                                                    (double2*)loss, pd);
        count_launch();
    }
    GP_CUDA(cudaGetLastError());
    return GP_OK;
}

// [rows] pieces of `width` bytes between pitched buffers: one contiguous copy when both
// pitches equal the width, else one 1-D copy per row (never a pitched 2-D copy).
static cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                             int rows, cudaMemcpyKind kind, cudaStream_t st) {
    if (rows <= 0 || width == 0) return cudaSuccess;
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, st);
    for (int r = 0; r < rows; ++r) {
        cudaError_t e = cudaMemcpyAsync((char*)dst + (size_t)r * dpitch, (const char*)src + (size_t)r * spitch,
                                        width, kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Host-buffer path: chunk the batch, pipeline H2D / solve / D2H over NSTREAM streams.
extern "C" int gp_fbs_solve_batch_host(GpFeeder* f, int64_t S, int64_t ld, const double* spu_h,
                                       double* V_h, double* I_h, double* Vmag_h, double* loss_h,
                                       int32_t* iters_h, int32_t* status_h, double* diff_h,
                                       double tol, int32_t maxiter) {
    int rc = fbs_check(f, S, ld, "gp_fbs_solve_batch_host");
    if (rc) return rc;
    if (S == 0) return GP_OK;
    FbsImpl* im = (FbsImpl*)f->impl;
    if (!spu_h && im->n_srows > 0) return fail(GP_EINVAL, "gp_fbs_solve_batch_host: null spu");
    GP_CUDA(cudaSetDevice(f->device));
    // The solve kernel is latency-bound for a single wave, so a chunk should be as
    // large as possible while still giving every stream work to overlap with the
    // other streams' copies: S/NSTREAM scenarios, at most ~64 MB of rows.
    const int64_t rows = (int64_t)im->n_vrows + im->n_irows + im->n_srows;
    // Rows are copied one 1-D transfer each (pitched 2-D copies run at ~7 GB/s on this
    // platform, 1-D pinned copies at ~55 GB/s), so a chunk should also keep a row piece
    // >= 256 KB (16 Ki scenarios) when memory allows: cap 3 GiB of rows per stream.
    // Measured on B200 (scripts/e2e_probe.py): every row piece is its own DMA (~2-4 us of
    // overhead), so a batch that moves <= 128 MB is fastest as ONE chunk (three contiguous
    // copies around the kernels); larger batches are cut into NSTREAM pipelined chunks with
    // row pieces of at least 512 KB (32 Ki scenarios).
    // One-launch path: with the feeder-specialised kernel and pinned (device-mapped) host buffers the
    // kernel reads spu and writes its results over PCIe itself, tile by tile: upload, solve and
    // download overlap inside one launch and there is nothing to chunk.
    // (a handle whose device kernel is the depth-first build writes its V rows every sweep, which must not cross
    // PCIe: its host path runs the two-sweep build, compiled here on first use, when that build fits at all)
    if (im->jit_kernel && (!im->jit_dfs || im->jit_alt_warps >= 1) && maxiter >= 1 &&
        g_host_zero_copy && (g_fbs_kernel == 0 || g_fbs_kernel == 5) &&
        ld * 16 < ((int64_t)1 << 32)) {
        const void* hp[] = {spu_h, V_h, I_h, Vmag_h, loss_h, iters_h, status_h, diff_h};
        void* dp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        bool ok = iters_h && status_h && diff_h;
        for (int i = 0; i < 8 && ok; ++i) {
            if (!hp[i]) continue;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, hp[i]) != cudaSuccess || at.type != cudaMemoryTypeHost ||
                !at.devicePointer) {
                cudaGetLastError();
                ok = false;
            } else {
                dp[i] = at.devicePointer;
            }
        }
        if (ok) {
            const bool host_ig = im->jit_dfs ? im->jit_alt_ig : im->jit_ig;     // I rows of the host path's build
            const size_t b_stage = ((size_t)im->n_srows * (size_t)ld * 16 + 255) & ~(size_t)255;
            const size_t need_stage = b_stage + (host_ig ? (size_t)im->n_irows * (size_t)ld * 16 : 0) + 256;
            if (need_stage > im->zc_bytes) {
                if (im->d_zc) cudaFree(im->d_zc);
                im->d_zc = nullptr;
                im->zc_bytes = 0;
                GP_CUDA(cudaMalloc(&im->d_zc, need_stage));
                im->zc_bytes = need_stage;
            }
            if (!im->streams[0]) GP_CUDA(cudaStreamCreateWithFlags(&im->streams[0], cudaStreamNonBlocking));
            // the host path's own build of the kernel (next tile's PCIe reads issued before this tile's stores);
            // feeders with more than 32 load rows, or a failed build, keep the device-path kernel
            const int dev_warps = im->jit_dfs ? im->jit_alt_warps : im->jit_warps;
            const int want_host_warps = dev_warps < g_host_warps ? dev_warps : g_host_warps;
            if (im->jit_host_tried && im->jit_kernel_host && im->jit_warps_host != want_host_warps) {
                // gp_set_option("host_warps") changed since this handle's host kernel was built: build again
                GP_CUDA(cudaStreamSynchronize(im->streams[0]));
                driver().ModuleUnload(im->jit_module_host);
                im->jit_module_host = im->jit_kernel_host = nullptr;
                im->jit_host_tried = false;
            }
            if (!im->jit_host_tried && ((im->n_srows <= 32 && g_host_read_ahead) || im->jit_dfs)) {
                im->jit_host_tried = true;
                im->jit_warps_host = want_host_warps;
                if (fbs_jit_build(f, im->jit_warps_host, host_ig, g_host_read_ahead != 0, &im->jit_module_host,
                                  &im->jit_kernel_host) != GP_OK)
                    im->jit_module_host = im->jit_kernel_host = nullptr;
            }
            if (im->jit_dfs && !im->jit_kernel_host) goto staged;       // no two-sweep build: copy - solve - copy
            // I rows kept in global memory must stay on the device; the host copy is a second destination
            double* I_state = host_ig ? (double*)((char*)im->d_zc + b_stage) : (double*)dp[2];
            rc = fbs_jit_launch(im, S, ld, (const double*)dp[0], (double*)dp[1], I_state, (int32_t*)dp[5],
                                (int32_t*)dp[6], (double*)dp[7], tol, maxiter, im->streams[0], (double*)dp[3],
                                (double*)dp[4], im->d_zc, host_ig ? (double*)dp[2] : nullptr, nullptr,
                                (g_host_read_ahead || im->jit_dfs) ? im->jit_kernel_host : nullptr, im->jit_warps_host,
                                nullptr, 0, im->n_vrows + (host_ig ? 0 : im->n_islots));
            if (rc) return rc;
            count_launch();
            GP_CUDA(cudaStreamSynchronize(im->streams[0]));
            return GP_OK;
        }
    }
staged:
    // (measured on B200, 123-bus, 140,160 scenarios, 0.8 GB over PCIe: 2 chunks 7.3e6 solves/s, 4: 6.8e6, 8: 5.6e6,
    // 16: 4.1e6 - the solve of a chunk is latency-bound and takes as long as the whole batch's, and every row piece
    // is its own DMA, so more and smaller chunks only add overhead)
    int want_chunks = g_host_chunks > 0 ? g_host_chunks : (rows * S * 16 <= ((int64_t)128 << 20) ? 1 : 2);
    int64_t C = ((S + want_chunks - 1) / want_chunks + 127) / 128 * 128;
    if (C < 32768 && g_host_chunks == 0) C = 32768;
    const int64_t cap = (((int64_t)3 << 30) / (rows * 16)) / 128 * 128;
    if (C > cap) C = cap;
    if (C < 1024) C = 1024;
    if (C > S) C = (S + 127) / 128 * 128;
    // per-chunk device layout: spu | V | I | Vmag | loss | diff | iters | status
    const size_t b_spu = (size_t)im->n_srows * C * 16, b_V = (size_t)im->n_vrows * C * 16,
                 b_I = (size_t)im->n_irows * C * 16, b_mag = (size_t)im->n_vrows * C * 8,
                 b_loss = (size_t)C * 16, b_diff = (size_t)C * 8, b_i32 = (size_t)C * 4;
    const size_t need = b_spu + b_V + b_I + b_mag + b_loss + b_diff + 2 * b_i32 + 1024;
    if (need > im->stage_bytes) {
        im->stage_bytes = 0;        // stays 0 unless every slot is allocated
        for (int i = 0; i < FbsImpl::NSTREAM; ++i) {
            if (im->d_stage[i]) cudaFree(im->d_stage[i]);
            im->d_stage[i] = nullptr;
            GP_CUDA(cudaMalloc(&im->d_stage[i], need));
            if (!im->streams[i]) GP_CUDA(cudaStreamCreateWithFlags(&im->streams[i], cudaStreamNonBlocking));
        }
        im->stage_bytes = need;
    }
    int slot = 0;
    for (int64_t s0 = 0; s0 < S; s0 += C, slot = (slot + 1) % FbsImpl::NSTREAM) {
        const int64_t n = (S - s0 < C) ? (S - s0) : C;
        cudaStream_t st = im->streams[slot];
        char* base = (char*)im->d_stage[slot];
        double* d_spu = (double*)base;
        double* d_V = (double*)(base + b_spu);
        double* d_I = (double*)(base + b_spu + b_V);
        double* d_mag = (double*)(base + b_spu + b_V + b_I);
        double* d_loss = (double*)(base + b_spu + b_V + b_I + b_mag);
        double* d_diff = (double*)(base + b_spu + b_V + b_I + b_mag + b_loss);
        int32_t* d_it = (int32_t*)(base + b_spu + b_V + b_I + b_mag + b_loss + b_diff);
        int32_t* d_st = d_it + C;
        if (im->n_srows > 0)
            GP_CUDA(copy_rows(d_spu, (size_t)C * 16, spu_h + 2 * s0, (size_t)ld * 16, (size_t)n * 16,
                              im->n_srows, cudaMemcpyHostToDevice, st));
        rc = gp_fbs_solve_post_batch(f, n, C, d_spu, d_V, d_I, d_it, d_st, d_diff, Vmag_h ? d_mag : nullptr,
                                     loss_h ? d_loss : nullptr, tol, maxiter, st);
        if (rc) return rc;
        if (V_h)
            GP_CUDA(copy_rows(V_h + 2 * s0, (size_t)ld * 16, d_V, (size_t)C * 16, (size_t)n * 16,
                              im->n_vrows, cudaMemcpyDeviceToHost, st));
        if (I_h && im->n_irows > 0)
            GP_CUDA(copy_rows(I_h + 2 * s0, (size_t)ld * 16, d_I, (size_t)C * 16, (size_t)n * 16,
                              im->n_irows, cudaMemcpyDeviceToHost, st));
        if (Vmag_h)
            GP_CUDA(copy_rows(Vmag_h + s0, (size_t)ld * 8, d_mag, (size_t)C * 8, (size_t)n * 8,
                              im->n_vrows, cudaMemcpyDeviceToHost, st));
        if (loss_h) GP_CUDA(cudaMemcpyAsync(loss_h + 2 * s0, d_loss, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        if (diff_h) GP_CUDA(cudaMemcpyAsync(diff_h + s0, d_diff, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (iters_h) GP_CUDA(cudaMemcpyAsync(iters_h + s0, d_it, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        if (status_h) GP_CUDA(cudaMemcpyAsync(status_h + s0, d_st, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < FbsImpl::NSTREAM; ++i)
        if (im->streams[i]) GP_CUDA(cudaStreamSynchronize(im->streams[i]));
    return GP_OK;
}
