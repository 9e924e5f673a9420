// Batched three-phase Newton-Raphson (NR3) for sm_100a.
//
// The reference assembles a dense Jacobian and solves inv(J'J) J' F
// (solution_nr3.py:618-627).  On a radial feeder J has the structure of the
// tree: per non-root bus k with incoming edge l from parent p
//
//   KVL_l (:141-158):  E dv_p - dv_k - Z di_l            = r1
//   KCL_k (:265-299):  D dv_k + M di_l - M sum_c di_c    = r2
//
// so the Newton step J dx = F is solved exactly by block elimination along the
// tree: a backward sweep (children first) that carries per bus a 6x6 equivalent
// admittance Y_k and a 6-vector a_k with  di_l = a_k - Y_k dv_p, then a forward
// sweep that applies dx.  Residual and Jacobian values are formed on the fly
// from X (closed forms of the reference's H, g, b tensors), fused with the
// elimination; nothing of size N x N is ever built.  The pivots are the KVL
// identity blocks and M = [[A,B],[B,-A]] (|det| = |V|^2 ~ 1), so the fixed
// elimination order needs no pivoting.
//
// One thread owns one scenario (FP64, whole Newton loop in one launch, each
// scenario stops at its own iteration).  State and the per-bus (Y, a, r1)
// fronts live in HBM as [row][ld] arrays with the scenario index innermost:
// every access of a warp is a coalesced 128-bit transaction.
#include "gp_common.cuh"

#include <new>
#include <vector>

namespace gp {

struct __align__(16) Nr3Step {
    int bus_row[3];
    int child_count;
    int par_row[3];
    int child_begin;
    int edge_row[3];    // current entering the bus (line / transformer current, regulator in-leg)
    int kind;           // GP_EDGE_LINE or GP_EDGE_VR
    int load_srow[3];
    int aux;            // regulator steps: front slot holding (Dt, u, r_a) for the forward sweep
    int out_row[3];     // current leaving the parent (= edge_row except the regulator out-leg)
    int pad0;
    double cap[3];
    double pad1;
    double gamma[3];
    double pad2;
    double2 z[9];
    int gam_row[3];     // row of the per-scenario gamma array (GpSolveExtras.gamma_dev) per regulated phase, else -1
    int pad3;
};
static_assert(sizeof(Nr3Step) == 80 + 32 + 32 + 144 + 16, "Nr3Step layout");

struct Nr3Const {
    int root_row[3];
    int n_steps;
    double2 vslack[3];
    double cA[3], cB[3];
    double betaS, aPQ, aI, aZ;
    double f0_absent;
};

struct Nr3Impl {
    int n_steps = 0, n_vrows = 0, n_irows = 0, n_srows = 0, n_child = 0, n_lines = 0;
    Nr3Step* d_steps = nullptr;
    int4* d_child = nullptr;      // {child step, child edge rows[3]}
    int* d_lines = nullptr;       // [n_lines][9] irow, tx vrow, rx vrow (X row indices)
    int* d_vrow_srow = nullptr;   // [n_vrows]
    double* d_vrow_cap = nullptr; // [n_vrows]
    Nr3Const k;
    int n_grows = 0;    // regulated (step, phase) pairs = rows of a per-scenario gamma array
    int n_fronts = 0;   // n_steps + one extra slot per regulator step
    bool lines_only = true;   // no regulator step: eligible for the phase-parallel kernel
    bool mostly_3ph = false;  // at least 5 of 6 non-root buses carry all three phases
};

constexpr int FRONT_D2 = 24;   // double2 per step: Y (18) + a (3) + r1 (3)

// In-place LU (Doolittle, no pivoting) of a 6x6 held in registers.
__device__ __forceinline__ void lu6(double (&T)[6][6]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double inv = 1.0 / T[k][k];
#pragma unroll
        for (int r = k + 1; r < 6; ++r) {
            const double m = T[r][k] * inv;
            T[r][k] = m;
#pragma unroll
            for (int c = k + 1; c < 6; ++c) T[r][c] = fma(-m, T[k][c], T[r][c]);
        }
        T[k][k] = inv;   // keep the reciprocal of the pivot
    }
}

__device__ __forceinline__ void lu6_solve(const double (&T)[6][6], double (&b)[6]) {
#pragma unroll
    for (int r = 1; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < r; ++c) b[r] = fma(-T[r][c], b[c], b[r]);
#pragma unroll
    for (int r = 5; r >= 0; --r) {
#pragma unroll
        for (int c = r + 1; c < 6; ++c) b[r] = fma(-T[r][c], b[c], b[r]);
        b[r] *= T[r][r];
    }
}

// Load that is never predicated or branched around: an absent row reads a hot dummy row instead
// and the value is discarded.  With a branch (or a predicate the compiler turns into one) per
// phase, each phase's loads were issued only after the previous phase's arithmetic - three
// exposed memory round trips per step instead of one.  The asm keeps the compiler from
// re-introducing the control flow; its memory clobber orders it after the thread's own stores.
__device__ __forceinline__ double2 ld_sel(bool have, const double2* p, const double2* dummy) {
    const double2* a = have ? p : dummy;
    double2 t;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(a) : "memory");
    return have ? t : make_double2(0.0, 0.0);
}

// EX = the instance behind gp_nr3_solve_batch_ex: per-scenario regulator ratios (gam != NULL) and / or a
// warm start (X holds an earlier solution, solution_nr3.py:598).  The plain instance has neither.
template <int BLOCK, bool EX>
__global__ void __launch_bounds__(BLOCK)
nr3_solve_kernel(const Nr3Step* __restrict__ steps, const int4* __restrict__ childs, Nr3Const k,
                 int n_vrows, int n_irows, long long S, long long ld,
                 const double2* __restrict__ spu, double2* __restrict__ X,
                 double2* __restrict__ W, double2* __restrict__ DV, int* __restrict__ iters_out,
                 int* __restrict__ status_out, double* __restrict__ fmax_out, int maxiter, int PF,
                 const double* __restrict__ gam, int warm, const double* __restrict__ wpu) {
    const long long s = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (s >= S) return;
    const double* Gs = (EX && gam) ? gam + s : nullptr;
    // per-scenario Solution.wpu rows [V row][ld]: change_KCL_matrices subtracts wpu from the constant term of BOTH
    // KCL rows of a bus-phase (solution_nr3.py:588)
    const double* Wp = (EX && wpu) ? wpu + s : nullptr;
    const bool flat = !(EX && warm);
    const double2* spu_s = spu + s;
    double2* Xs = X + s;
    double2* Ws = W + s;
    double2* DVs = DV + s;
    const int n_steps = k.n_steps;
    const double2* dummy = Xs + (long long)k.root_row[0] * ld;
    // L2 prefetch of the rows a step reads that were written a sweep ago (X rows and loads in the
    // backward sweep, the step's front and X rows in the forward sweep), PF steps ahead.  With
    // about one warp per SM sub-partition (BASELINE config C3) every one of those loads is an
    // exposed DRAM round trip otherwise.
    auto pf = [&](const double2* base, int row) {
        if (row >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)row * ld));
    };

    // flat start (solution_nr3.py:37-82): every bus-phase at VSLACK, currents 0.
    // Absent phases are not stored: their rows of J are identity rows whose first
    // Newton step sets them to exactly 0 (:417-419, 159-161).
    for (int pos = 0; flat && pos < n_steps; ++pos) {
        const int4 bv = __ldg(reinterpret_cast<const int4*>(steps[pos].bus_row));
        const int4 ev = __ldg(reinterpret_cast<const int4*>(steps[pos].edge_row));
        const int4 ov = __ldg(reinterpret_cast<const int4*>(steps[pos].out_row));
        const int brow[3] = {bv.x, bv.y, bv.z}, erow[3] = {ev.x, ev.y, ev.z}, orow[3] = {ov.x, ov.y, ov.z};
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (brow[p] >= 0) Xs[(long long)brow[p] * ld] = k.vslack[p];
            if (erow[p] >= 0) Xs[(long long)erow[p] * ld] = make_double2(0.0, 0.0);
            if (orow[p] >= 0) Xs[(long long)orow[p] * ld] = make_double2(0.0, 0.0);
        }
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) Xs[(long long)k.root_row[p] * ld] = k.vslack[p];

    int it = 0;
    double fm = 1e99;
    bool nanflag = false;
    while (fm >= 1e-9 && !nanflag && it < maxiter) {        // :618
        // (absent phases sit at VSLACK only in the flat start; an earlier solution has them at 0)
        fm = (it == 0 && flat) ? k.f0_absent : 0.0;
        // ================= backward sweep: residual, Jacobian blocks, elimination =========
        for (int pos = n_steps - 1; pos >= 0; --pos) {
            if (PF > 0 && pos >= PF) {
                const Nr3Step* sn = steps + (pos - PF);
                const int4 b2 = __ldg(reinterpret_cast<const int4*>(sn->bus_row));
                const int4 e2 = __ldg(reinterpret_cast<const int4*>(sn->edge_row));
                const int4 l2 = __ldg(reinterpret_cast<const int4*>(sn->load_srow));
                pf(Xs, b2.x); pf(Xs, b2.y); pf(Xs, b2.z);
                pf(Xs, e2.x); pf(Xs, e2.y); pf(Xs, e2.z);
                pf(spu_s, l2.x); pf(spu_s, l2.y); pf(spu_s, l2.z);
            }
            const Nr3Step* st = steps + pos;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_row));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_row));
            const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_row));
            const int4 lv = __ldg(reinterpret_cast<const int4*>(st->load_srow));
            const int brow[3] = {bv.x, bv.y, bv.z}, prow[3] = {pv.x, pv.y, pv.z};
            const int erow[3] = {ev.x, ev.y, ev.z}, lrow[3] = {lv.x, lv.y, lv.z};
            const int child_count = bv.w, child_begin = pv.w, kind = ev.w, aux = lv.w;
            // ---- single-phase bus on a Line/Transformer edge (55 of the 125 steps of the IEEE-123
            // feeder): the 6x6 system is the identity outside one 2x2 block, so the same elimination
            // (same operations, same order on that block) runs on 2-vectors: 7 loads instead of 12-36,
            // a 2x2 instead of a 6x6 LU and three instead of seven triangular solves.
            const int nph = (bv.x >= 0) + (bv.y >= 0) + (bv.z >= 0);
            if (nph == 1 && kind == GP_EDGE_LINE) {
                const int p = bv.x >= 0 ? 0 : (bv.y >= 0 ? 1 : 2);
                const int rb = p == 0 ? bv.x : (p == 1 ? bv.y : bv.z), re = p == 0 ? ev.x : (p == 1 ? ev.y : ev.z);
                const int rp = p == 0 ? pv.x : (p == 1 ? pv.y : pv.z), rl = p == 0 ? lv.x : (p == 1 ? lv.y : lv.z);
                const double2 v1 = ld_sel(true, Xs + (long long)rb * ld, dummy);
                double2 in1 = ld_sel(true, Xs + (long long)re * ld, dummy);
                const double2 vp1 = ld_sel(rp >= 0, Xs + (long long)rp * ld, dummy);
                const double2 sp1 = ld_sel(rl >= 0, spu_s + (long long)rl * ld, dummy);
                const double2 z = __ldg(&st->z[4 * p]);                 // Z[p][p]
                const double cap = __ldg(&st->cap[p]);
                double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0, ra = 0.0, rb2 = 0.0;     // Dt block, rt
                // KVL residual (:141-158)
                double2 r1s = make_double2(vp1.x - v1.x, vp1.y - v1.y);
                r1s.x -= z.x * in1.x - z.y * in1.y;
                r1s.y -= z.x * in1.y + z.y * in1.x;
                nanflag |= (r1s.x != r1s.x) | (r1s.y != r1s.y);
                fm = fmax(fm, fmax(fabs(r1s.x), fabs(r1s.y)));
                for (int c = 0; c < child_count; ++c) {                 // children carry the same single phase
                    const int4 ch = __ldg(childs + child_begin + c);
                    const int cr = p == 0 ? ch.y : (p == 1 ? ch.z : ch.w);
                    const double2* Wc = Ws + (long long)ch.x * FRONT_D2 * ld;
                    const bool hc = cr >= 0;
                    const double2 ic = ld_sel(hc, Xs + (long long)cr * ld, dummy);
                    const double2 ac = ld_sel(hc, Wc + (long long)(18 + p) * ld, dummy);
                    const double2 y0c = ld_sel(hc, Wc + (long long)((2 * p) * 3 + p) * ld, dummy);
                    const double2 y1c = ld_sel(hc, Wc + (long long)((2 * p + 1) * 3 + p) * ld, dummy);
                    in1.x -= ic.x;
                    in1.y -= ic.y;
                    ra += ac.x;
                    rb2 += ac.y;
                    d00 += y0c.x;
                    d01 += y0c.y;
                    d10 += y1c.x;
                    d11 += y1c.y;
                }
                {   // KCL residual and D block (closed form of H, g, b; :226-299, 403-440)
                    const double A = v1.x, B = v1.y, Cn = in1.x, Dn = in1.y;
                    const double pl = sp1.x, ql = sp1.y;
                    const double cA = p == 0 ? k.cA[0] : (p == 1 ? k.cA[1] : k.cA[2]);
                    const double cB = p == 0 ? k.cB[0] : (p == 1 ? k.cB[1] : k.cB[2]);
                    const double wq = (EX && Wp) ? __ldg(Wp + (long long)rb * ld) : 0.0;
                    const double hAr = -pl * cA, hBr = -pl * cB, b0r = -pl * k.betaS - wq;
                    const double hAi = -ql * cA + cap * cA;
                    const double hBi = -ql * cB + cap * cB;
                    const double b0i = (-ql * k.betaS) + (cap * k.betaS) - wq;
                    const double fr = A * Cn + B * Dn + hAr * A * A + hBr * B * B + b0r;
                    const double fi = B * Cn - A * Dn + hAi * A * A + hBi * B * B + b0i;
                    nanflag |= (fr != fr) | (fi != fi);
                    fm = fmax(fm, fmax(fabs(fr), fabs(fi)));
                    const double e00 = Cn + 2.0 * hAr * A, e01 = Dn + 2.0 * hBr * B;
                    const double e10 = -Dn + 2.0 * hAi * A, e11 = Cn + 2.0 * hBi * B;
                    const double inv = 1.0 / (A * A + B * B);
                    d00 += (A * e00 + B * e10) * inv;
                    d01 += (A * e01 + B * e11) * inv;
                    d10 += (B * e00 - A * e10) * inv;
                    d11 += (B * e01 - A * e11) * inv;
                    ra += (A * fr + B * fi) * inv;
                    rb2 += (B * fr - A * fi) * inv;
                }
                // rhs = rt + Dt r1 ;  T = I - Dt Zr (Zr = [[R, -X], [X, R]])
                ra = fma(d00, r1s.x, ra);
                ra = fma(d01, r1s.y, ra);
                rb2 = fma(d10, r1s.x, rb2);
                rb2 = fma(d11, r1s.y, rb2);
                const double t00 = 1.0 - (d00 * z.x + d01 * z.y), t01 = 0.0 - (-d00 * z.y + d01 * z.x);
                const double t10 = 0.0 - (d10 * z.x + d11 * z.y), t11 = 1.0 - (-d10 * z.y + d11 * z.x);
                // Doolittle LU of the 2x2 block, as lu6 does it
                const double i0 = 1.0 / t00, mlu = t10 * i0;
                const double i1 = 1.0 / fma(-mlu, t01, t11);
                // a = T^-1 rhs ; Y = T^-1 Dt (two columns)
                double xa = ra, xb = fma(-mlu, ra, rb2);
                xb *= i1;
                xa = fma(-t01, xb, xa) * i0;
                double y0a = d00, y0b = fma(-mlu, d00, d10);
                y0b *= i1;
                y0a = fma(-t01, y0b, y0a) * i0;
                double y1a = d01, y1b = fma(-mlu, d01, d11);
                y1b *= i1;
                y1a = fma(-t01, y1b, y1a) * i0;
                double2* Wk1 = Ws + (long long)pos * FRONT_D2 * ld;
                Wk1[(long long)(18 + p) * ld] = make_double2(xa, xb);
                Wk1[(long long)(21 + p) * ld] = r1s;
                Wk1[(long long)((2 * p) * 3 + p) * ld] = make_double2(y0a, y1a);
                Wk1[(long long)((2 * p + 1) * 3 + p) * ld] = make_double2(y0b, y1b);
                continue;
            }
            // every load of the step is issued before the first use (one exposed round trip, not one per phase)
            // (each load is its own predicated instruction - no branch around a group of them - so that
            // the scheduler can put all of them in flight together)
            const double2 zz = make_double2(0.0, 0.0);
            double2 v[3], inet[3], r1[3], vpar[3], spl[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const bool hb = brow[p] >= 0;
                r1[p] = zz;
                v[p] = ld_sel(hb, Xs + (long long)brow[p] * ld, dummy);
                inet[p] = ld_sel(hb, Xs + (long long)erow[p] * ld, dummy);
                vpar[p] = ld_sel(hb && prow[p] >= 0, Xs + (long long)prow[p] * ld, dummy);
                spl[p] = ld_sel(hb && lrow[p] >= 0, spu_s + (long long)lrow[p] * ld, dummy);
            }
            if (kind == GP_EDGE_LINE) {
                // KVL residual r1 = E v_p - v - Z i   (:141-158; transformers :165-175 with Z = 0)
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    if (brow[p] < 0) continue;
                    r1[p].x = vpar[p].x - v[p].x;
                    r1[p].y = vpar[p].y - v[p].y;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (brow[q] < 0) continue;
                        const double2 z = __ldg(&st->z[3 * p + q]);
                        r1[p].x -= z.x * inet[q].x - z.y * inet[q].y;
                        r1[p].y -= z.x * inet[q].y + z.y * inet[q].x;
                    }
                    nanflag |= (r1[p].x != r1[p].x) | (r1[p].y != r1[p].y);
                    fm = fmax(fm, fmax(fabs(r1[p].x), fabs(r1[p].y)));
                }
            }
            double2 iin[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) iin[p] = inet[p];
            double Dt[6][6], rt[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                rt[r] = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) Dt[r][c] = 0.0;
            }
            // children: net current, and their fronts (Y_c, a_c)
            for (int c = 0; c < child_count; ++c) {
                const int4 ch = __ldg(childs + child_begin + c);
                const int crow[3] = {ch.y, ch.z, ch.w};
                const double2* Wc = Ws + (long long)ch.x * FRONT_D2 * ld;
                double2 ic[3], ac[3], yc[6][3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {           // issue the child's whole front, then consume it
                    const bool hp = crow[p] >= 0;
                    ic[p] = ld_sel(hp, Xs + (long long)crow[p] * ld, dummy);
                    ac[p] = ld_sel(hp, Wc + (long long)(18 + p) * ld, dummy);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {       // only the child's own phases were stored
                        const bool hq = hp && crow[q] >= 0;
                        yc[2 * p][q] = ld_sel(hq, Wc + (long long)((2 * p) * 3 + q) * ld, dummy);
                        yc[2 * p + 1][q] = ld_sel(hq, Wc + (long long)((2 * p + 1) * 3 + q) * ld, dummy);
                    }
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    inet[p].x -= ic[p].x;
                    inet[p].y -= ic[p].y;
                    rt[2 * p] += ac[p].x;
                    rt[2 * p + 1] += ac[p].y;
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        Dt[2 * p][2 * q] += yc[2 * p][q].x;
                        Dt[2 * p][2 * q + 1] += yc[2 * p][q].y;
                        Dt[2 * p + 1][2 * q] += yc[2 * p + 1][q].x;
                        Dt[2 * p + 1][2 * q + 1] += yc[2 * p + 1][q].y;
                    }
                }
            }
            // KCL residual and D block per phase (closed form of H, g, b; :226-299, 403-440)
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if (brow[p] < 0) continue;
                const double A = v[p].x, B = v[p].y, Cn = inet[p].x, Dn = inet[p].y;
                const double pl = spl[p].x, ql = spl[p].y;
                const double cap = __ldg(&st->cap[p]);
                const double wq = (EX && Wp) ? __ldg(Wp + (long long)brow[p] * ld) : 0.0;
                const double hAr = -pl * k.cA[p], hBr = -pl * k.cB[p], b0r = -pl * k.betaS - wq;
                const double hAi = -ql * k.cA[p] + cap * k.cA[p];
                const double hBi = -ql * k.cB[p] + cap * k.cB[p];
                const double b0i = (-ql * k.betaS) + (cap * k.betaS) - wq;
                const double fr = A * Cn + B * Dn + hAr * A * A + hBr * B * B + b0r;
                const double fi = B * Cn - A * Dn + hAi * A * A + hBi * B * B + b0i;
                nanflag |= (fr != fr) | (fi != fi);
                fm = fmax(fm, fmax(fabs(fr), fabs(fi)));
                const double d00 = Cn + 2.0 * hAr * A, d01 = Dn + 2.0 * hBr * B;
                const double d10 = -Dn + 2.0 * hAi * A, d11 = Cn + 2.0 * hBi * B;
                const double inv = 1.0 / (A * A + B * B);
                // M^-1 = M / |V|^2,  M = [[A, B], [B, -A]]
                Dt[2 * p][2 * p] += (A * d00 + B * d10) * inv;
                Dt[2 * p][2 * p + 1] += (A * d01 + B * d11) * inv;
                Dt[2 * p + 1][2 * p] += (B * d00 - A * d10) * inv;
                Dt[2 * p + 1][2 * p + 1] += (B * d01 - A * d11) * inv;
                rt[2 * p] += (A * fr + B * fi) * inv;
                rt[2 * p + 1] += (B * fr - A * fi) * inv;
            }
            double2* Wk = Ws + (long long)pos * FRONT_D2 * ld;
            if (kind == GP_EDGE_VR) {
                // ---- regulator edge (:444-515): v_r = gamma v_m ; V_m conj(I_out) = V_r conj(I_in).
                //   dv_r = gamma dv_m + r_a,  di_in = u - Dt gamma dv_m  (u = rt - Dt r_a),
                //   di_out = a_k - Y_k dv_m  with the 2x2 blocks M(v) = [[A,B],[B,-A]], N(i) = [[C,D],[-D,C]]:
                //   a_k = Mm^-1 [F/2 + N_in r_a + M_r u],  Y_k = Mm^-1 [N_out - g N_in + M_r Dt g]
                const int4 ov = __ldg(reinterpret_cast<const int4*>(st->out_row));
                const int orow[3] = {ov.x, ov.y, ov.z};
                double2 vm[3], io[3], ra[3];
                double g[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    vm[p] = io[p] = ra[p] = make_double2(0.0, 0.0);
                    g[p] = 0.0;
                    if (brow[p] < 0) continue;
                    g[p] = __ldg(&st->gamma[p]);
                    if (EX && Gs) g[p] = __ldg(Gs + (long long)__ldg(&st->gam_row[p]) * ld);
                    vm[p] = vpar[p];
                    io[p] = Xs[(long long)orow[p] * ld];
                    ra[p] = make_double2(-g[p] * vm[p].x + v[p].x, -g[p] * vm[p].y + v[p].y);
                    const double f0 = 2.0 * ((vm[p].x * io[p].x + vm[p].y * io[p].y) - (v[p].x * iin[p].x + v[p].y * iin[p].y));
                    const double f1 = 2.0 * ((vm[p].y * io[p].x - vm[p].x * io[p].y) - (v[p].y * iin[p].x - v[p].x * iin[p].y));
                    nanflag |= (f0 != f0) | (f1 != f1) | (ra[p].x != ra[p].x) | (ra[p].y != ra[p].y);
                    fm = fmax(fm, fmax(fmax(fabs(f0), fabs(f1)), fmax(fabs(ra[p].x), fabs(ra[p].y))));
                    r1[p] = make_double2(0.5 * f0, 0.5 * f1);       // reuse r1 as F/2
                }
                double u[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    u[r] = rt[r];
#pragma unroll
                    for (int q = 0; q < 3; ++q) u[r] -= Dt[r][2 * q] * ra[q].x + Dt[r][2 * q + 1] * ra[q].y;
                }
                double2* Wa = Ws + (long long)aux * FRONT_D2 * ld;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    if (brow[p] < 0) continue;
                    Wa[(long long)(18 + p) * ld] = make_double2(u[2 * p], u[2 * p + 1]);
                    Wa[(long long)(21 + p) * ld] = ra[p];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (brow[q] < 0) continue;
                        Wa[(long long)((2 * p) * 3 + q) * ld] = make_double2(Dt[2 * p][2 * q], Dt[2 * p][2 * q + 1]);
                        Wa[(long long)((2 * p + 1) * 3 + q) * ld] = make_double2(Dt[2 * p + 1][2 * q], Dt[2 * p + 1][2 * q + 1]);
                    }
                    const double A1 = vm[p].x, B1 = vm[p].y, A2 = v[p].x, B2 = v[p].y;
                    const double Co = io[p].x, Do = io[p].y, Ci = iin[p].x, Di = iin[p].y;
                    const double im = 1.0 / (A1 * A1 + B1 * B1);
                    // a_k rows of this phase
                    const double t0 = r1[p].x + (Ci * ra[p].x + Di * ra[p].y) + (A2 * u[2 * p] + B2 * u[2 * p + 1]);
                    const double t1 = r1[p].y + (-Di * ra[p].x + Ci * ra[p].y) + (B2 * u[2 * p] - A2 * u[2 * p + 1]);
                    Wk[(long long)(18 + p) * ld] = make_double2((A1 * t0 + B1 * t1) * im, (B1 * t0 - A1 * t1) * im);
                    // Y_k rows of this phase: Mm^-1 [ M_r (Dt[rows p] * g) + (N_out - g N_in) on the diagonal block ]
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (brow[q] < 0) continue;
                        double w00 = (A2 * Dt[2 * p][2 * q] + B2 * Dt[2 * p + 1][2 * q]) * g[q];
                        double w01 = (A2 * Dt[2 * p][2 * q + 1] + B2 * Dt[2 * p + 1][2 * q + 1]) * g[q];
                        double w10 = (B2 * Dt[2 * p][2 * q] - A2 * Dt[2 * p + 1][2 * q]) * g[q];
                        double w11 = (B2 * Dt[2 * p][2 * q + 1] - A2 * Dt[2 * p + 1][2 * q + 1]) * g[q];
                        if (q == p) {
                            w00 += Co - g[p] * Ci;
                            w01 += Do - g[p] * Di;
                            w10 += -Do + g[p] * Di;
                            w11 += Co - g[p] * Ci;
                        }
                        Wk[(long long)((2 * p) * 3 + q) * ld] = make_double2((A1 * w00 + B1 * w10) * im, (A1 * w01 + B1 * w11) * im);
                        Wk[(long long)((2 * p + 1) * 3 + q) * ld] = make_double2((B1 * w00 - A1 * w10) * im, (B1 * w01 - A1 * w11) * im);
                    }
                }
                continue;
            }
            // rhs_a = rt + Dt r1 ;  T = I - Dt Zr
            double T[6][6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    rt[r] = fma(Dt[r][2 * q], r1[q].x, rt[r]);
                    rt[r] = fma(Dt[r][2 * q + 1], r1[q].y, rt[r]);
                    double t0 = (r == 2 * q) ? 1.0 : 0.0, t1 = (r == 2 * q + 1) ? 1.0 : 0.0;
                    if (brow[q] >= 0) {
#pragma unroll
                        for (int w = 0; w < 3; ++w) {
                            if (brow[w] < 0) continue;
                            const double2 z = __ldg(&st->z[3 * w + q]);   // Z[w][q]
                            // Zr[2w][2q] = R, Zr[2w][2q+1] = -X, Zr[2w+1][2q] = X, Zr[2w+1][2q+1] = R
                            t0 -= Dt[r][2 * w] * z.x + Dt[r][2 * w + 1] * z.y;
                            t1 -= -Dt[r][2 * w] * z.y + Dt[r][2 * w + 1] * z.x;
                        }
                    }
                    T[r][2 * q] = t0;
                    T[r][2 * q + 1] = t1;
                }
            }
            lu6(T);
            lu6_solve(T, rt);
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if (brow[p] < 0) continue;
                Wk[(long long)(18 + p) * ld] = make_double2(rt[2 * p], rt[2 * p + 1]);
                Wk[(long long)(21 + p) * ld] = r1[p];
            }
            // Y = T^-1 Dt E, two columns (one phase of the parent) at a time
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (brow[q] < 0) continue;
                double y0[6], y1[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    y0[r] = Dt[r][2 * q];
                    y1[r] = Dt[r][2 * q + 1];
                }
                lu6_solve(T, y0);
                lu6_solve(T, y1);
#pragma unroll
                for (int r = 0; r < 6; ++r)
                    if (brow[r / 2] >= 0) Wk[(long long)(r * 3 + q) * ld] = make_double2(y0[r], y1[r]);
            }
        }
        // ================= slack rows (:84-109) and forward sweep ================================
        double2 dv_prev[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const double2 v0 = Xs[(long long)k.root_row[p] * ld];
            const double2 d = make_double2(v0.x - k.vslack[p].x, v0.y - k.vslack[p].y);
            nanflag |= (d.x != d.x) | (d.y != d.y);
            fm = fmax(fm, fmax(fabs(d.x), fabs(d.y)));
            DVs[(long long)k.root_row[p] * ld] = d;
            Xs[(long long)k.root_row[p] * ld] = make_double2(v0.x - d.x, v0.y - d.y);
            dv_prev[p] = d;
        }
        // The forward sweep runs in reverse post-order: a bus is followed immediately by the child
        // visited first, so on every chain the parent's dv is the one just computed: it is taken
        // from registers (dv_prev, rows prev_row) instead of a store -> load round trip through L2.
        int prev_row[3] = {k.root_row[0], k.root_row[1], k.root_row[2]};
        for (int pos = 0; pos < n_steps; ++pos) {
            if (PF > 0 && pos + PF < n_steps) {
                const Nr3Step* sn = steps + (pos + PF);
                const int4 b2 = __ldg(reinterpret_cast<const int4*>(sn->bus_row));
                const int4 e2 = __ldg(reinterpret_cast<const int4*>(sn->edge_row));
                const int br2[3] = {b2.x, b2.y, b2.z};
                pf(Xs, b2.x); pf(Xs, b2.y); pf(Xs, b2.z);
                pf(Xs, e2.x); pf(Xs, e2.y); pf(Xs, e2.z);
                const double2* Wn = Ws + (long long)(pos + PF) * FRONT_D2 * ld;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    if (br2[p] < 0) continue;
                    pf(Wn, 18 + p);
                    pf(Wn, 21 + p);
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        if (br2[q] >= 0) {
                            pf(Wn, (2 * p) * 3 + q);
                            pf(Wn, (2 * p + 1) * 3 + q);
                        }
                }
            }
            const Nr3Step* st = steps + pos;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_row));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_row));
            const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_row));
            const int brow[3] = {bv.x, bv.y, bv.z}, prow[3] = {pv.x, pv.y, pv.z};
            const int erow[3] = {ev.x, ev.y, ev.z};
            const int kind = ev.w;
            const double2* Wk = Ws + (long long)pos * FRONT_D2 * ld;
            const double2 zz = make_double2(0.0, 0.0);
            if ((bv.x >= 0) + (bv.y >= 0) + (bv.z >= 0) == 1 && kind == GP_EDGE_LINE) {   // single-phase bus
                const int p = bv.x >= 0 ? 0 : (bv.y >= 0 ? 1 : 2);
                const int rb = p == 0 ? bv.x : (p == 1 ? bv.y : bv.z), re = p == 0 ? ev.x : (p == 1 ? ev.y : ev.z);
                const int rp = p == 0 ? pv.x : (p == 1 ? pv.y : pv.z);
                const int rprev = p == 0 ? prev_row[0] : (p == 1 ? prev_row[1] : prev_row[2]);
                const double2 dprev = p == 0 ? dv_prev[0] : (p == 1 ? dv_prev[1] : dv_prev[2]);
                const bool hpar = rp >= 0, chain = rp == rprev;
                const double2 dg = ld_sel(hpar && !chain, DVs + (long long)rp * ld, dummy);
                const double2 dvp1 = (hpar && chain) ? dprev : dg;
                double2 di1 = ld_sel(true, Wk + (long long)(18 + p) * ld, dummy);
                const double2 r1f1 = ld_sel(true, Wk + (long long)(21 + p) * ld, dummy);
                const double2 xv1 = ld_sel(true, Xs + (long long)rb * ld, dummy);
                const double2 xi1 = ld_sel(true, Xs + (long long)re * ld, dummy);
                const double2 y0 = ld_sel(true, Wk + (long long)((2 * p) * 3 + p) * ld, dummy);
                const double2 y1 = ld_sel(true, Wk + (long long)((2 * p + 1) * 3 + p) * ld, dummy);
                const double2 z = __ldg(&st->z[4 * p]);
                di1.x -= y0.x * dvp1.x + y0.y * dvp1.y;
                di1.y -= y1.x * dvp1.x + y1.y * dvp1.y;
                double2 dv1 = make_double2(dvp1.x - r1f1.x, dvp1.y - r1f1.y);
                dv1.x -= z.x * di1.x - z.y * di1.y;
                dv1.y -= z.x * di1.y + z.y * di1.x;
                Xs[(long long)rb * ld] = make_double2(xv1.x - dv1.x, xv1.y - dv1.y);
                Xs[(long long)re * ld] = make_double2(xi1.x - di1.x, xi1.y - di1.y);
                DVs[(long long)rb * ld] = dv1;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    prev_row[q] = q == p ? rb : -1;
                    dv_prev[q] = dv1;
                }
                continue;
            }
            double2 dvp[3], di[3], dv[3], yf[6][3], r1f[3], xv[3], xi[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {               // issue every load of the step first (predicated, no branches)
                const bool hb = brow[p] >= 0;
                const bool hpar = hb && prow[p] >= 0, chain = prow[p] == prev_row[p];
                const double2 dg = ld_sel(hpar && !chain, DVs + (long long)prow[p] * ld, dummy);
                dvp[p] = (hpar && chain) ? dv_prev[p] : dg;
                di[p] = ld_sel(hb, Wk + (long long)(18 + p) * ld, dummy);
                r1f[p] = ld_sel(hb, Wk + (long long)(21 + p) * ld, dummy);
                xv[p] = ld_sel(hb, Xs + (long long)brow[p] * ld, dummy);
                xi[p] = ld_sel(hb, Xs + (long long)erow[p] * ld, dummy);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const bool hq = hb && brow[q] >= 0;
                    yf[2 * p][q] = ld_sel(hq, Wk + (long long)((2 * p) * 3 + q) * ld, dummy);
                    yf[2 * p + 1][q] = ld_sel(hq, Wk + (long long)((2 * p + 1) * 3 + q) * ld, dummy);
                }
            }
            // current leaving the parent: di = a - Y dv_p
#pragma unroll
            for (int p = 0; p < 3; ++p) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    di[p].x -= yf[2 * p][q].x * dvp[q].x + yf[2 * p][q].y * dvp[q].y;
                    di[p].y -= yf[2 * p + 1][q].x * dvp[q].x + yf[2 * p + 1][q].y * dvp[q].y;
                }
            }
            if (kind == GP_EDGE_VR) {
                const int4 lv = __ldg(reinterpret_cast<const int4*>(st->load_srow));
                const int4 ov = __ldg(reinterpret_cast<const int4*>(st->out_row));
                const int orow[3] = {ov.x, ov.y, ov.z};
                const double2* Wa = Ws + (long long)lv.w * FRONT_D2 * ld;
                double2 gdv[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    double g = __ldg(&st->gamma[p]);
                    if (EX && Gs && brow[p] >= 0) g = __ldg(Gs + (long long)__ldg(&st->gam_row[p]) * ld);
                    gdv[p] = make_double2(g * dvp[p].x, g * dvp[p].y);
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    if (brow[p] < 0) continue;
                    const double2 ra = Wa[(long long)(21 + p) * ld];
                    double2 dii = Wa[(long long)(18 + p) * ld];          // u
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (brow[q] < 0) continue;
                        const double2 d0 = Wa[(long long)((2 * p) * 3 + q) * ld];
                        const double2 d1 = Wa[(long long)((2 * p + 1) * 3 + q) * ld];
                        dii.x -= d0.x * gdv[q].x + d0.y * gdv[q].y;
                        dii.y -= d1.x * gdv[q].x + d1.y * gdv[q].y;
                    }
                    dv[p] = make_double2(gdv[p].x + ra.x, gdv[p].y + ra.y);
                    const double2 xo = Xs[(long long)orow[p] * ld];
                    Xs[(long long)brow[p] * ld] = make_double2(xv[p].x - dv[p].x, xv[p].y - dv[p].y);
                    Xs[(long long)erow[p] * ld] = make_double2(xi[p].x - dii.x, xi[p].y - dii.y);
                    Xs[(long long)orow[p] * ld] = make_double2(xo.x - di[p].x, xo.y - di[p].y);
                    DVs[(long long)brow[p] * ld] = dv[p];
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    prev_row[p] = brow[p];
                    dv_prev[p] = dv[p];
                }
                continue;
            }
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if (brow[p] < 0) continue;
                dv[p] = make_double2(dvp[p].x - r1f[p].x, dvp[p].y - r1f[p].y);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (brow[q] < 0) continue;
                    const double2 z = __ldg(&st->z[3 * p + q]);
                    dv[p].x -= z.x * di[q].x - z.y * di[q].y;
                    dv[p].y -= z.x * di[q].y + z.y * di[q].x;
                }
                Xs[(long long)brow[p] * ld] = make_double2(xv[p].x - dv[p].x, xv[p].y - dv[p].y);
                Xs[(long long)erow[p] * ld] = make_double2(xi[p].x - di[p].x, xi[p].y - di[p].y);
                DVs[(long long)brow[p] * ld] = dv[p];
            }
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                prev_row[p] = brow[p];
                dv_prev[p] = dv[p];
            }
        }
        ++it;
    }
    iters_out[s] = it;
    fmax_out[s] = nanflag ? __longlong_as_double(0x7ff8000000000000LL) : fm;
    status_out[s] = nanflag ? GP_ST_NONFINITE : (fm < 1e-9 ? GP_ST_CONVERGED : (isinf(fm) ? GP_ST_NONFINITE : GP_ST_MAXITER));
}

// ---------------------------------------------------------------------------------------
// Phase-parallel Newton kernel (Line / Transformer feeders).
//
// A scenario is owned by THREE lanes, one per phase (lane = phase * 10 + j: ten scenarios per
// warp, lanes 30 and 31 idle).  Lane g holds the two real rows 2g, 2g+1 of everything the step
// builds - Dt, the right-hand sides, T = I - Dt Zr - and touches only the memory rows of its own
// phase.  The 6x6 system is reduced by Gauss-Jordan on the augmented rows [T | rhs | Dt]: per
// pivot the owning lane's row is broadcast with warp shuffles and every lane eliminates the
// pivot column from its own two rows, so the seven solves of the thread-per-scenario kernel
// (a = T^-1 rhs, Y = T^-1 Dt) come out of one sweep over the pivots with no back-substitution.
// Compared with one thread per scenario this needs about 130 registers instead of 254 (twice
// the resident warps), a third of the dependent chain per step, and three times the warps for
// a given batch - which is what BASELINE's small NR3 batches (16,384 scenarios = one warp per
// SM sub-partition at a thread per scenario) lack.  Same equations, same stop rule.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double2 shfl_d2(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ int pick3(const int4 v, int g) { return g == 0 ? v.x : (g == 1 ? v.y : v.z); }
__device__ __forceinline__ int pick3c(const int4 v, int g) { return g == 0 ? v.y : (g == 1 ? v.z : v.w); }

// RING: the rows of a step whose addresses do not depend on this sweep's arithmetic (backward: the bus's X rows and
// load; forward: its front, r1 and X rows) are staged two steps ahead in shared memory with cp.async - the kernel's
// registers (150, 12 warps per SM, all of which the BASELINE batch sizes need) leave no room to stage them there,
// its shared memory is otherwise unused.  A lane copies and reads only its own 16-byte column: no barrier.
template <int BLOCK, bool RING>
__global__ void __launch_bounds__(BLOCK, 3)     // <= 168 registers: 12 warps per SM, what C3 (11.1) and C4-NR3 need
nr3_phase_kernel(const Nr3Step* __restrict__ steps, const int4* __restrict__ childs, Nr3Const k,
                 long long S, long long ld, const double2* __restrict__ spu, double2* __restrict__ X,
                 double2* __restrict__ W, double2* __restrict__ DV, int* __restrict__ iters_out,
                 int* __restrict__ status_out, double* __restrict__ fmax_out, int maxiter, int PF) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * BLOCK + threadIdx.x) >> 5;
    const int j = lane % 10;
    const int g = lane < 30 ? lane / 10 : 2;
    const long long s_raw = warp * 10 + j;
    const bool owner = lane < 30 && s_raw < S;
    const long long s = s_raw < S ? s_raw : S - 1;          // idle lanes shadow a real scenario (loads only)
    const int src0 = j, src1 = 10 + j, src2 = 20 + j;
    const double2* spu_s = spu + s;
    double2* Xs = X + s;
    double2* Ws = W + s;
    double2* DVs = DV + s;
    const int n_steps = k.n_steps;
    const double2 zz = make_double2(0.0, 0.0);
    const double2* dummy = Xs + (long long)k.root_row[0] * ld;
    const int my_root = g == 0 ? k.root_row[0] : (g == 1 ? k.root_row[1] : k.root_row[2]);
    const double2 my_slack = g == 0 ? k.vslack[0] : (g == 1 ? k.vslack[1] : k.vslack[2]);
    const double my_cA = g == 0 ? k.cA[0] : (g == 1 ? k.cA[1] : k.cA[2]);
    const double my_cB = g == 0 ? k.cB[0] : (g == 1 ? k.cB[1] : k.cB[2]);
    // L2 prefetch, PF steps ahead, of the rows of this lane's phase that were written a sweep ago: X rows and the load
    // in the backward sweep, the step's front and X rows in the forward sweep.  At the BASELINE batch sizes a scenario's
    // state is far beyond its share of L2 (C3: 20 kB per scenario, 330 MB in all), so each of them is a DRAM round trip
    // that the three warps per scheduler this kernel's registers allow cannot hide.
    auto pf = [&](const double2* base, long long row) {
        if (row >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + row * ld));
    };
    constexpr int NSTAGE = 3, NSLOT = 10;               // [stage][slot][lane] per warp, 15 kB
    extern __shared__ __align__(16) double2 nr3_ring[];
    double2* const my_ring = nr3_ring + (size_t)(threadIdx.x >> 5) * (NSTAGE * NSLOT * 32) + lane;
    auto cpa = [&](int stage, int slot, const double2* src, bool have) {     // absent rows are zero-filled
        const unsigned dst = (unsigned)__cvta_generic_to_shared(my_ring + (stage * NSLOT + slot) * 32);
        const int n = have ? 16 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(have ? src : dummy), "r"(n) : "memory");
    };
    auto staged = [&](int stage, int slot) { return my_ring[(stage * NSLOT + slot) * 32]; };
    auto issue_bwd = [&](int pos) {
        if (pos >= 0) {
            const Nr3Step* sn = steps + pos;
            const int b = pick3(__ldg(reinterpret_cast<const int4*>(sn->bus_row)), g);
            const int e = pick3(__ldg(reinterpret_cast<const int4*>(sn->edge_row)), g);
            const int pr = pick3(__ldg(reinterpret_cast<const int4*>(sn->par_row)), g);
            const int l = pick3(__ldg(reinterpret_cast<const int4*>(sn->load_srow)), g);
            const int stg = pos % NSTAGE;
            cpa(stg, 0, Xs + (long long)b * ld, b >= 0);
            cpa(stg, 1, Xs + (long long)e * ld, b >= 0 && e >= 0);
            cpa(stg, 2, Xs + (long long)pr * ld, b >= 0 && pr >= 0);
            cpa(stg, 3, spu_s + (long long)l * ld, b >= 0 && l >= 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto issue_fwd = [&](int pos) {
        if (pos < n_steps) {
            const Nr3Step* sn = steps + pos;
            const int4 b4 = __ldg(reinterpret_cast<const int4*>(sn->bus_row));
            const int b = pick3(b4, g);
            const int e = pick3(__ldg(reinterpret_cast<const int4*>(sn->edge_row)), g);
            const double2* Wn = Ws + (long long)pos * FRONT_D2 * ld;
            const int stg = pos % NSTAGE;
            const bool hb2 = b >= 0;
            cpa(stg, 0, Wn + (long long)(18 + g) * ld, hb2);
            cpa(stg, 1, Wn + (long long)(21 + g) * ld, hb2);
            cpa(stg, 2, Xs + (long long)b * ld, hb2);
            cpa(stg, 3, Xs + (long long)e * ld, hb2 && e >= 0);
            cpa(stg, 4, Wn + (long long)((2 * g) * 3 + 0) * ld, hb2 && b4.x >= 0);
            cpa(stg, 5, Wn + (long long)((2 * g) * 3 + 1) * ld, hb2 && b4.y >= 0);
            cpa(stg, 6, Wn + (long long)((2 * g) * 3 + 2) * ld, hb2 && b4.z >= 0);
            cpa(stg, 7, Wn + (long long)((2 * g + 1) * 3 + 0) * ld, hb2 && b4.x >= 0);
            cpa(stg, 8, Wn + (long long)((2 * g + 1) * 3 + 1) * ld, hb2 && b4.y >= 0);
            cpa(stg, 9, Wn + (long long)((2 * g + 1) * 3 + 2) * ld, hb2 && b4.z >= 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // flat start (solution_nr3.py:37-82)
    if (owner) {
        for (int pos = 0; pos < n_steps; ++pos) {
            const int b = pick3(__ldg(reinterpret_cast<const int4*>(steps[pos].bus_row)), g);
            const int e = pick3(__ldg(reinterpret_cast<const int4*>(steps[pos].edge_row)), g);
            if (b >= 0) Xs[(long long)b * ld] = my_slack;
            if (e >= 0) Xs[(long long)e * ld] = zz;
        }
        Xs[(long long)my_root * ld] = my_slack;
    }
    __syncwarp();

    int it = 0;
    double fm = 1e99;
    bool nanflag = false;
    bool active = owner && maxiter > 0;
    while (__any_sync(0xffffffffu, active)) {
        double fml = (it == 0) ? k.f0_absent : 0.0;          // this lane's share of max |F|
        bool nanl = false;
        // ================= backward sweep =================
        if (RING) {
            issue_bwd(n_steps - 1);
            issue_bwd(n_steps - 2);
        }
        for (int pos = n_steps - 1; pos >= 0; --pos) {
            if (RING) {
                issue_bwd(pos - 2);
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            }
            if (!RING && PF > 0 && pos >= PF) {
                const Nr3Step* sn = steps + (pos - PF);
                const int b2 = pick3(__ldg(reinterpret_cast<const int4*>(sn->bus_row)), g);
                if (b2 >= 0) {
                    pf(Xs, b2);
                    pf(Xs, pick3(__ldg(reinterpret_cast<const int4*>(sn->edge_row)), g));
                    pf(spu_s, pick3(__ldg(reinterpret_cast<const int4*>(sn->load_srow)), g));
                }
            }
            const Nr3Step* st = steps + pos;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_row));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_row));
            const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_row));
            const int4 lv = __ldg(reinterpret_cast<const int4*>(st->load_srow));
            const int my_b = pick3(bv, g), my_p = pick3(pv, g), my_e = pick3(ev, g), my_l = pick3(lv, g);
            const bool hb = my_b >= 0;
            const int child_count = bv.w, child_begin = pv.w;
            if ((bv.x >= 0) + (bv.y >= 0) + (bv.z >= 0) == 1) {
                // single-phase bus: the system is the identity outside one 2x2 block, which the lane of
                // that phase solves alone (no shuffles); the other two lanes have nothing to do
                if (hb) {
                    const int stg = pos % NSTAGE;
                    const double2 v1 = RING ? staged(stg, 0) : ld_sel(true, Xs + (long long)my_b * ld, dummy);
                    double2 in1 = RING ? staged(stg, 1) : ld_sel(true, Xs + (long long)my_e * ld, dummy);
                    const double2 vp1 = RING ? staged(stg, 2) : ld_sel(my_p >= 0, Xs + (long long)my_p * ld, dummy);
                    const double2 sp1 = RING ? staged(stg, 3) : ld_sel(my_l >= 0, spu_s + (long long)my_l * ld, dummy);
                    const double2 z = g == 0 ? __ldg(&st->z[0]) : (g == 1 ? __ldg(&st->z[4]) : __ldg(&st->z[8]));
                    const double cap = g == 0 ? __ldg(&st->cap[0]) : (g == 1 ? __ldg(&st->cap[1]) : __ldg(&st->cap[2]));
                    double d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0, ra = 0.0, rb2 = 0.0;
                    double2 r1s = make_double2(vp1.x - v1.x, vp1.y - v1.y);
                    r1s.x -= z.x * in1.x - z.y * in1.y;
                    r1s.y -= z.x * in1.y + z.y * in1.x;
                    nanl |= (r1s.x != r1s.x) | (r1s.y != r1s.y);
                    fml = fmax(fml, fmax(fabs(r1s.x), fabs(r1s.y)));
                    for (int c = 0; c < child_count; ++c) {
                        const int4 ch = __ldg(childs + child_begin + c);
                        const int cr = pick3c(ch, g);
                        const double2* Wc = Ws + (long long)ch.x * FRONT_D2 * ld;
                        const bool hc = cr >= 0;
                        const double2 ic = ld_sel(hc, Xs + (long long)cr * ld, dummy);
                        const double2 ac = ld_sel(hc, Wc + (long long)(18 + g) * ld, dummy);
                        const double2 y0c = ld_sel(hc, Wc + (long long)((2 * g) * 3 + g) * ld, dummy);
                        const double2 y1c = ld_sel(hc, Wc + (long long)((2 * g + 1) * 3 + g) * ld, dummy);
                        in1.x -= ic.x;
                        in1.y -= ic.y;
                        ra += ac.x;
                        rb2 += ac.y;
                        d00 += y0c.x;
                        d01 += y0c.y;
                        d10 += y1c.x;
                        d11 += y1c.y;
                    }
                    {
                        const double A = v1.x, B = v1.y, Cn = in1.x, Dn = in1.y;
                        const double pl = sp1.x, ql = sp1.y;
                        const double hAr = -pl * my_cA, hBr = -pl * my_cB, b0r = -pl * k.betaS;
                        const double hAi = -ql * my_cA + cap * my_cA;
                        const double hBi = -ql * my_cB + cap * my_cB;
                        const double b0i = (-ql * k.betaS) + (cap * k.betaS);
                        const double fr = A * Cn + B * Dn + hAr * A * A + hBr * B * B + b0r;
                        const double fi = B * Cn - A * Dn + hAi * A * A + hBi * B * B + b0i;
                        nanl |= (fr != fr) | (fi != fi);
                        fml = fmax(fml, fmax(fabs(fr), fabs(fi)));
                        const double e00 = Cn + 2.0 * hAr * A, e01 = Dn + 2.0 * hBr * B;
                        const double e10 = -Dn + 2.0 * hAi * A, e11 = Cn + 2.0 * hBi * B;
                        const double inv = 1.0 / (A * A + B * B);
                        d00 += (A * e00 + B * e10) * inv;
                        d01 += (A * e01 + B * e11) * inv;
                        d10 += (B * e00 - A * e10) * inv;
                        d11 += (B * e01 - A * e11) * inv;
                        ra += (A * fr + B * fi) * inv;
                        rb2 += (B * fr - A * fi) * inv;
                    }
                    ra = fma(d00, r1s.x, ra);
                    ra = fma(d01, r1s.y, ra);
                    rb2 = fma(d10, r1s.x, rb2);
                    rb2 = fma(d11, r1s.y, rb2);
                    const double t00 = 1.0 - (d00 * z.x + d01 * z.y), t01 = 0.0 - (-d00 * z.y + d01 * z.x);
                    const double t10 = 0.0 - (d10 * z.x + d11 * z.y), t11 = 1.0 - (-d10 * z.y + d11 * z.x);
                    const double i0 = 1.0 / t00, mlu = t10 * i0;
                    const double i1 = 1.0 / fma(-mlu, t01, t11);
                    double xa = ra, xb = fma(-mlu, ra, rb2);
                    xb *= i1;
                    xa = fma(-t01, xb, xa) * i0;
                    double y0a = d00, y0b = fma(-mlu, d00, d10);
                    y0b *= i1;
                    y0a = fma(-t01, y0b, y0a) * i0;
                    double y1a = d01, y1b = fma(-mlu, d01, d11);
                    y1b *= i1;
                    y1a = fma(-t01, y1b, y1a) * i0;
                    if (active) {
                        double2* Wk1 = Ws + (long long)pos * FRONT_D2 * ld;
                        Wk1[(long long)(18 + g) * ld] = make_double2(xa, xb);
                        Wk1[(long long)(21 + g) * ld] = r1s;
                        Wk1[(long long)((2 * g) * 3 + g) * ld] = make_double2(y0a, y1a);
                        Wk1[(long long)((2 * g + 1) * 3 + g) * ld] = make_double2(y0b, y1b);
                    }
                }
                continue;
            }
            const int stg = pos % NSTAGE;
            const double2 v = RING ? staged(stg, 0) : ld_sel(hb, Xs + (long long)my_b * ld, dummy);
            double2 inet = RING ? staged(stg, 1) : ld_sel(hb, Xs + (long long)my_e * ld, dummy);
            const double2 vpar = RING ? staged(stg, 2) : ld_sel(hb && my_p >= 0, Xs + (long long)my_p * ld, dummy);
            const double2 spl = RING ? staged(stg, 3) : ld_sel(hb && my_l >= 0, spu_s + (long long)my_l * ld, dummy);
            // KVL residual r1 = E v_p - v - Z i (:141-158): the currents of all three phases
            const double2 i0 = shfl_d2(inet, src0), i1 = shfl_d2(inet, src1), i2 = shfl_d2(inet, src2);
            const double2 zg0 = __ldg(&st->z[3 * g]), zg1 = __ldg(&st->z[3 * g + 1]), zg2 = __ldg(&st->z[3 * g + 2]);
            double2 r1 = zz;
            if (hb) {
                r1.x = vpar.x - v.x;
                r1.y = vpar.y - v.y;
                r1.x -= zg0.x * i0.x - zg0.y * i0.y;
                r1.y -= zg0.x * i0.y + zg0.y * i0.x;
                r1.x -= zg1.x * i1.x - zg1.y * i1.y;
                r1.y -= zg1.x * i1.y + zg1.y * i1.x;
                r1.x -= zg2.x * i2.x - zg2.y * i2.y;
                r1.y -= zg2.x * i2.y + zg2.y * i2.x;
                nanl |= (r1.x != r1.x) | (r1.y != r1.y);
                fml = fmax(fml, fmax(fabs(r1.x), fabs(r1.y)));
            }
            // augmented rows of this lane: columns 0-5 T, 6 rhs, 7-12 Dt
            double R0[13], R1[13];
#pragma unroll
            for (int c = 0; c < 13; ++c) R0[c] = R1[c] = 0.0;
            for (int c = 0; c < child_count; ++c) {         // children: net current, fronts (Y_c, a_c)
                const int4 ch = __ldg(childs + child_begin + c);
                const int cr = pick3c(ch, g);
                const bool hc = cr >= 0;
                const double2* Wc = Ws + (long long)ch.x * FRONT_D2 * ld;
                const double2 ic = ld_sel(hc, Xs + (long long)cr * ld, dummy);
                const double2 ac = ld_sel(hc, Wc + (long long)(18 + g) * ld, dummy);
                double2 y0[3], y1[3];
                const bool hq[3] = {hc && ch.y >= 0, hc && ch.z >= 0, hc && ch.w >= 0};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    y0[q] = ld_sel(hq[q], Wc + (long long)((2 * g) * 3 + q) * ld, dummy);
                    y1[q] = ld_sel(hq[q], Wc + (long long)((2 * g + 1) * 3 + q) * ld, dummy);
                }
                inet.x -= ic.x;
                inet.y -= ic.y;
                R0[6] += ac.x;
                R1[6] += ac.y;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    R0[7 + 2 * q] += y0[q].x;
                    R0[8 + 2 * q] += y0[q].y;
                    R1[7 + 2 * q] += y1[q].x;
                    R1[8 + 2 * q] += y1[q].y;
                }
            }
            if (hb) {       // KCL residual and D block of this phase (:226-299, 403-440)
                const double A = v.x, B = v.y, Cn = inet.x, Dn = inet.y;
                const double pl = spl.x, ql = spl.y;
                const double cap = g == 0 ? __ldg(&st->cap[0]) : (g == 1 ? __ldg(&st->cap[1]) : __ldg(&st->cap[2]));
                const double hAr = -pl * my_cA, hBr = -pl * my_cB, b0r = -pl * k.betaS;
                const double hAi = -ql * my_cA + cap * my_cA;
                const double hBi = -ql * my_cB + cap * my_cB;
                const double b0i = (-ql * k.betaS) + (cap * k.betaS);
                const double fr = A * Cn + B * Dn + hAr * A * A + hBr * B * B + b0r;
                const double fi = B * Cn - A * Dn + hAi * A * A + hBi * B * B + b0i;
                nanl |= (fr != fr) | (fi != fi);
                fml = fmax(fml, fmax(fabs(fr), fabs(fi)));
                const double d00 = Cn + 2.0 * hAr * A, d01 = Dn + 2.0 * hBr * B;
                const double d10 = -Dn + 2.0 * hAi * A, d11 = Cn + 2.0 * hBi * B;
                const double inv = 1.0 / (A * A + B * B);
                const double e00 = (A * d00 + B * d10) * inv, e01 = (A * d01 + B * d11) * inv;
                const double e10 = (B * d00 - A * d10) * inv, e11 = (B * d01 - A * d11) * inv;
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (q == g) {
                        R0[7 + 2 * q] += e00;
                        R0[8 + 2 * q] += e01;
                        R1[7 + 2 * q] += e10;
                        R1[8 + 2 * q] += e11;
                    }
                R0[6] += (A * fr + B * fi) * inv;
                R1[6] += (B * fr - A * fi) * inv;
            }
            // rhs = rt + Dt r1 ;  T = I - Dt Zr
            {
                const double2 ra = shfl_d2(r1, src0), rb = shfl_d2(r1, src1), rc = shfl_d2(r1, src2);
                R0[6] = fma(R0[7], ra.x, R0[6]);
                R0[6] = fma(R0[8], ra.y, R0[6]);
                R0[6] = fma(R0[9], rb.x, R0[6]);
                R0[6] = fma(R0[10], rb.y, R0[6]);
                R0[6] = fma(R0[11], rc.x, R0[6]);
                R0[6] = fma(R0[12], rc.y, R0[6]);
                R1[6] = fma(R1[7], ra.x, R1[6]);
                R1[6] = fma(R1[8], ra.y, R1[6]);
                R1[6] = fma(R1[9], rb.x, R1[6]);
                R1[6] = fma(R1[10], rb.y, R1[6]);
                R1[6] = fma(R1[11], rc.x, R1[6]);
                R1[6] = fma(R1[12], rc.y, R1[6]);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                double t00 = (q == g) ? 1.0 : 0.0, t01 = 0.0, t10 = 0.0, t11 = (q == g) ? 1.0 : 0.0;
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const double2 z = __ldg(&st->z[3 * w + q]);      // Z[w][q]; zero where a phase is absent
                    t00 -= R0[7 + 2 * w] * z.x + R0[8 + 2 * w] * z.y;
                    t01 -= -R0[7 + 2 * w] * z.y + R0[8 + 2 * w] * z.x;
                    t10 -= R1[7 + 2 * w] * z.x + R1[8 + 2 * w] * z.y;
                    t11 -= -R1[7 + 2 * w] * z.y + R1[8 + 2 * w] * z.x;
                }
                R0[2 * q] = t00;
                R0[2 * q + 1] = t01;
                R1[2 * q] = t10;
                R1[2 * q + 1] = t11;
            }
            // Gauss-Jordan over the six pivots; pivot rows stay unnormalised, 1/pivot is kept per row
            double inv0 = 1.0, inv1 = 1.0;
#pragma unroll
            for (int kk = 0; kk < 6; ++kk) {
                const int own = (kk >> 1) * 10 + j;                  // lane that owns pivot row kk
                const bool mine = (kk >> 1) == g;
                double pvr[13];
#pragma unroll
                for (int c = kk; c < 13; ++c) pvr[c] = __shfl_sync(0xffffffffu, (kk & 1) ? R1[c] : R0[c], own);
                const double pinv = 1.0 / pvr[kk];
                if (mine) {
                    if (kk & 1) inv1 = pinv;
                    else inv0 = pinv;
                }
                const double f0 = (mine && !(kk & 1)) ? 0.0 : R0[kk] * pinv;
                const double f1 = (mine && (kk & 1)) ? 0.0 : R1[kk] * pinv;
#pragma unroll
                for (int c = kk + 1; c < 13; ++c) {
                    R0[c] = fma(-f0, pvr[c], R0[c]);
                    R1[c] = fma(-f1, pvr[c], R1[c]);
                }
            }
            if (hb && active) {
                double2* Wk = Ws + (long long)pos * FRONT_D2 * ld;
                Wk[(long long)(18 + g) * ld] = make_double2(R0[6] * inv0, R1[6] * inv1);
                Wk[(long long)(21 + g) * ld] = r1;
                const int br[3] = {bv.x, bv.y, bv.z};
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    if (br[q] >= 0) {
                        Wk[(long long)((2 * g) * 3 + q) * ld] = make_double2(R0[7 + 2 * q] * inv0, R0[8 + 2 * q] * inv0);
                        Wk[(long long)((2 * g + 1) * 3 + q) * ld] = make_double2(R1[7 + 2 * q] * inv1, R1[8 + 2 * q] * inv1);
                    }
            }
        }
        // ================= slack rows (:84-109), stop rule, forward sweep =================
        {
            const double2 v0 = Xs[(long long)my_root * ld];
            const double2 d = make_double2(v0.x - my_slack.x, v0.y - my_slack.y);
            nanl |= (d.x != d.x) | (d.y != d.y);
            fml = fmax(fml, fmax(fabs(d.x), fabs(d.y)));
            if (active) {
                DVs[(long long)my_root * ld] = d;
                Xs[(long long)my_root * ld] = make_double2(v0.x - d.x, v0.y - d.y);
            }
        }
        {   // max |F| and the NaN flag over the three phase lanes of the scenario
            const double f1 = __shfl_sync(0xffffffffu, fml, src0), f2 = __shfl_sync(0xffffffffu, fml, src1),
                         f3 = __shfl_sync(0xffffffffu, fml, src2);
            const unsigned nb = __ballot_sync(0xffffffffu, nanl);
            if (active) {
                fm = fmax(f1, fmax(f2, f3));
                nanflag = ((nb >> src0) | (nb >> src1) | (nb >> src2)) & 1u;
            }
        }
        __syncwarp();       // the dv of a parent is read by the other two lanes of the scenario below
        if (RING) {
            issue_fwd(0);
            issue_fwd(1);
        }
        for (int pos = 0; pos < n_steps; ++pos) {
            if (RING) {
                issue_fwd(pos + 2);
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            }
            if (!RING && PF > 0 && pos + PF < n_steps) {
                const Nr3Step* sn = steps + (pos + PF);
                const int4 b2 = __ldg(reinterpret_cast<const int4*>(sn->bus_row));
                if (pick3(b2, g) >= 0) {
                    const double2* Wn = Ws + (long long)(pos + PF) * FRONT_D2 * ld;
                    pf(Xs, pick3(b2, g));
                    pf(Xs, pick3(__ldg(reinterpret_cast<const int4*>(sn->edge_row)), g));
                    pf(Wn, 18 + g);
                    pf(Wn, 21 + g);
                    if (b2.x >= 0) { pf(Wn, (2 * g) * 3 + 0); pf(Wn, (2 * g + 1) * 3 + 0); }
                    if (b2.y >= 0) { pf(Wn, (2 * g) * 3 + 1); pf(Wn, (2 * g + 1) * 3 + 1); }
                    if (b2.z >= 0) { pf(Wn, (2 * g) * 3 + 2); pf(Wn, (2 * g + 1) * 3 + 2); }
                }
            }
            const Nr3Step* st = steps + pos;
            const int4 bv = __ldg(reinterpret_cast<const int4*>(st->bus_row));
            const int4 pv = __ldg(reinterpret_cast<const int4*>(st->par_row));
            const int4 ev = __ldg(reinterpret_cast<const int4*>(st->edge_row));
            const int my_b = pick3(bv, g), my_e = pick3(ev, g);
            const bool hb = my_b >= 0;
            const int br[3] = {bv.x, bv.y, bv.z}, pr[3] = {pv.x, pv.y, pv.z};
            const double2* Wk = Ws + (long long)pos * FRONT_D2 * ld;
            if ((bv.x >= 0) + (bv.y >= 0) + (bv.z >= 0) == 1) {       // single-phase bus: its lane alone
                if (hb) {
                    const int my_p = pick3(pv, g);
                    const double2 dvp1 = ld_sel(my_p >= 0, DVs + (long long)my_p * ld, dummy);
                    const int stg = pos % NSTAGE;
                    double2 di1 = RING ? staged(stg, 0) : ld_sel(true, Wk + (long long)(18 + g) * ld, dummy);
                    const double2 r1f1 = RING ? staged(stg, 1) : ld_sel(true, Wk + (long long)(21 + g) * ld, dummy);
                    const double2 xv1 = RING ? staged(stg, 2) : ld_sel(true, Xs + (long long)my_b * ld, dummy);
                    const double2 xi1 = RING ? staged(stg, 3) : ld_sel(true, Xs + (long long)my_e * ld, dummy);
                    const double2 ya = RING ? staged(stg, 4 + g) : ld_sel(true, Wk + (long long)((2 * g) * 3 + g) * ld, dummy);
                    const double2 yb = RING ? staged(stg, 7 + g) : ld_sel(true, Wk + (long long)((2 * g + 1) * 3 + g) * ld, dummy);
                    const double2 z = g == 0 ? __ldg(&st->z[0]) : (g == 1 ? __ldg(&st->z[4]) : __ldg(&st->z[8]));
                    di1.x -= ya.x * dvp1.x + ya.y * dvp1.y;
                    di1.y -= yb.x * dvp1.x + yb.y * dvp1.y;
                    double2 dv1 = make_double2(dvp1.x - r1f1.x, dvp1.y - r1f1.y);
                    dv1.x -= z.x * di1.x - z.y * di1.y;
                    dv1.y -= z.x * di1.y + z.y * di1.x;
                    if (active) {
                        Xs[(long long)my_b * ld] = make_double2(xv1.x - dv1.x, xv1.y - dv1.y);
                        Xs[(long long)my_e * ld] = make_double2(xi1.x - di1.x, xi1.y - di1.y);
                        DVs[(long long)my_b * ld] = dv1;
                    }
                }
                __syncwarp();
                continue;
            }
            const int stg = pos % NSTAGE;
            double2 di = RING ? staged(stg, 0) : ld_sel(hb, Wk + (long long)(18 + g) * ld, dummy);
            const double2 r1f = RING ? staged(stg, 1) : ld_sel(hb, Wk + (long long)(21 + g) * ld, dummy);
            const double2 xv = RING ? staged(stg, 2) : ld_sel(hb, Xs + (long long)my_b * ld, dummy);
            const double2 xi = RING ? staged(stg, 3) : ld_sel(hb, Xs + (long long)my_e * ld, dummy);
            double2 dvp[3], y0[3], y1[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const bool hq = hb && br[q] >= 0;
                dvp[q] = ld_sel(hq && pr[q] >= 0, DVs + (long long)pr[q] * ld, dummy);
                y0[q] = RING ? staged(stg, 4 + q) : ld_sel(hq, Wk + (long long)((2 * g) * 3 + q) * ld, dummy);
                y1[q] = RING ? staged(stg, 7 + q) : ld_sel(hq, Wk + (long long)((2 * g + 1) * 3 + q) * ld, dummy);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {                   // di = a - Y dv_p
                di.x -= y0[q].x * dvp[q].x + y0[q].y * dvp[q].y;
                di.y -= y1[q].x * dvp[q].x + y1[q].y * dvp[q].y;
            }
            const double2 d0 = shfl_d2(di, src0), d1 = shfl_d2(di, src1), d2 = shfl_d2(di, src2);
            if (hb && active) {
                const double2 zg0 = __ldg(&st->z[3 * g]), zg1 = __ldg(&st->z[3 * g + 1]), zg2 = __ldg(&st->z[3 * g + 2]);
                const double2 dp = g == 0 ? dvp[0] : (g == 1 ? dvp[1] : dvp[2]);
                double2 dv = make_double2(dp.x - r1f.x, dp.y - r1f.y);
                dv.x -= zg0.x * d0.x - zg0.y * d0.y;
                dv.y -= zg0.x * d0.y + zg0.y * d0.x;
                dv.x -= zg1.x * d1.x - zg1.y * d1.y;
                dv.y -= zg1.x * d1.y + zg1.y * d1.x;
                dv.x -= zg2.x * d2.x - zg2.y * d2.y;
                dv.y -= zg2.x * d2.y + zg2.y * d2.x;
                Xs[(long long)my_b * ld] = make_double2(xv.x - dv.x, xv.y - dv.y);
                Xs[(long long)my_e * ld] = make_double2(xi.x - di.x, xi.y - di.y);
                DVs[(long long)my_b * ld] = dv;
            }
            __syncwarp();   // dv rows written above are the dvp of the children, read by all three lanes
        }
        if (active) {
            ++it;
            active = fm >= 1e-9 && !nanflag && it < maxiter;        // :618
        }
    }
    if (owner && g == 0) {
        iters_out[s] = it;
        fmax_out[s] = nanflag ? __longlong_as_double(0x7ff8000000000000LL) : fm;
        status_out[s] = nanflag ? GP_ST_NONFINITE : (fm < 1e-9 ? GP_ST_CONVERGED : (isinf(fm) ? GP_ST_NONFINITE : GP_ST_MAXITER));
    }
}

// nr3_lib/map_output.py:24-27: |re| <= 1e-12 -> v = imag + 0j (imag lands in the REAL
// slot); else |im| <= 1e-12 -> v = re + 0j.
__device__ __forceinline__ double2 scrub(double2 v) {
    if (fabs(v.x) <= 1e-12) return make_double2(v.y, 0.0);
    if (fabs(v.y) <= 1e-12) return make_double2(v.x, 0.0);
    return v;
}

// X -> V, |V|, sV, i_Node per V row (map_output.py:18-29, 56-77)
__global__ void nr3_map_rows_kernel(int n_vrows, int n_irows, const int* __restrict__ vrow_srow,
                                    const double* __restrict__ vrow_cap, Nr3Const k, long long S,
                                    long long ld, const double2* __restrict__ spu,
                                    const double2* __restrict__ X, double2* __restrict__ V,
                                    double2* __restrict__ I, double2* __restrict__ sV,
                                    double2* __restrict__ iN, double* __restrict__ Vmag,
                                    const double* __restrict__ wpu) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int r = blockIdx.y; r < n_vrows + n_irows; r += gridDim.y) {
        const double2 x = scrub(X[(long long)r * ld + s]);
        if (r >= n_vrows) {
            if (I) I[(long long)(r - n_vrows) * ld + s] = x;
            continue;
        }
        if (V) V[(long long)r * ld + s] = x;
        const double a = sqrt(fma(x.x, x.x, x.y * x.y));
        if (Vmag) Vmag[(long long)r * ld + s] = a;
        if (sV || iN) {
            double2 sv = make_double2(0.0, 0.0);
            const int lr = __ldg(vrow_srow + r);
            if (lr >= 0) {
                const double2 sp = spu[(long long)lr * ld + s];
                const double f = k.aPQ + k.aI * a + k.aZ * (a * a);
                sv.x = sp.x * f;
                sv.y = sp.y * f;
            }
            sv.y -= __ldg(vrow_cap + r);          // no |V|^2 factor here (map_output.py:60)
            if (wpu) sv.y += wpu[(long long)r * ld + s];    // + 1j * wpu (map_output.py:60)
            sv = scrub(sv);
            if (sV) sV[(long long)r * ld + s] = sv;
            if (iN) {
                const double inv = 1.0 / fma(x.x, x.x, x.y * x.y);
                const double qr = (sv.x * x.x + sv.y * x.y) * inv;
                const double qi = (sv.y * x.x - sv.x * x.y) * inv;
                iN[(long long)r * ld + s] = scrub(make_double2(qr, -qi));
            }
        }
    }
}

__global__ void nr3_map_lines_kernel(int n_lines, int n_vrows, const int* __restrict__ lines, long long S,
                                     long long ld, const double2* __restrict__ X,
                                     double2* __restrict__ Stx, double2* __restrict__ Srx,
                                     double2* __restrict__ loss) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double2 acc = make_double2(0.0, 0.0);
    for (int l = 0; l < n_lines; ++l) {
        const int* rec = lines + 9 * l;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int ir = __ldg(rec + p);
            if (ir < 0) continue;
            const int tr = __ldg(rec + 3 + p), rr = __ldg(rec + 6 + p);
            const double2 i = scrub(X[(long long)ir * ld + s]);
            const double2 vt = tr >= 0 ? scrub(X[(long long)tr * ld + s]) : make_double2(0.0, 0.0);
            const double2 vr = rr >= 0 ? scrub(X[(long long)rr * ld + s]) : make_double2(0.0, 0.0);
            const double2 st = scrub(make_double2(vt.x * i.x + vt.y * i.y, vt.y * i.x - vt.x * i.y));
            const double2 sr = scrub(make_double2(vr.x * i.x + vr.y * i.y, vr.y * i.x - vr.x * i.y));
            if (Stx) Stx[(long long)(ir - n_vrows) * ld + s] = st;
            if (Srx) Srx[(long long)(ir - n_vrows) * ld + s] = sr;
            acc.x += st.x - sr.x;
            acc.y += st.y - sr.y;
        }
    }
    if (loss) loss[s] = acc;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_nr3_create(const GpNr3Desc* d, int device, GpFeeder** out) {
    if (!d || !out) return fail(GP_EINVAL, "gp_nr3_create: null argument");
    if (d->n_steps < 0 || d->n_vrows < 3 || d->n_irows < 0 || d->n_srows < 0)
        return fail(GP_EINVAL, "gp_nr3_create: bad sizes");
    GP_CUDA(cudaSetDevice(device));
    Nr3Impl* im = new (std::nothrow) Nr3Impl();
    if (!im) return fail(GP_ENOMEM, "gp_nr3_create: out of host memory");
    im->n_steps = d->n_steps;
    im->n_vrows = d->n_vrows;
    im->n_irows = d->n_irows;
    im->n_srows = d->n_srows;
    im->n_child = d->n_child;
    im->n_lines = d->n_lines;
    const int nx = d->n_vrows + d->n_irows;
    im->n_fronts = d->n_steps;
    std::vector<Nr3Step> steps((size_t)d->n_steps);
    std::vector<int> vrow_srow((size_t)d->n_vrows, -1);
    std::vector<double> vrow_cap((size_t)d->n_vrows, 0.0);
    for (int i = 0; i < d->n_steps; ++i) {
        Nr3Step& s = steps[i];
        memset(&s, 0, sizeof(s));
        s.child_begin = d->child_begin[i];
        s.child_count = d->child_count[i];
        s.kind = d->kind[i];
        s.aux = -1;
        if (s.kind == GP_EDGE_VR) {
            s.aux = im->n_fronts++;
            im->lines_only = false;
        }
        else if (s.kind != GP_EDGE_LINE) {
            delete im;
            return fail(GP_EINVAL, "gp_nr3_create: step %d: unknown edge kind %d", i, s.kind);
        }
        if (s.child_begin < 0 || s.child_count < 0 || s.child_begin + s.child_count > d->n_child) {
            delete im;
            return fail(GP_EINVAL, "gp_nr3_create: step %d child range out of bounds", i);
        }
        for (int p = 0; p < 3; ++p) {
            s.bus_row[p] = d->bus_vrow[3 * i + p];
            s.par_row[p] = d->par_vrow[3 * i + p];
            s.edge_row[p] = d->edge_irow[3 * i + p] >= 0 ? d->edge_irow[3 * i + p] + d->n_vrows : -1;
            s.out_row[p] = d->edge_orow[3 * i + p] >= 0 ? d->edge_orow[3 * i + p] + d->n_vrows : -1;
            s.gamma[p] = d->vr_gamma[3 * i + p];
            s.load_srow[p] = d->load_srow[3 * i + p];
            const bool has_b = s.bus_row[p] >= 0, has_e = s.edge_row[p] >= 0;
            if (has_b != (s.out_row[p] >= 0) || s.out_row[p] >= nx ||
                (s.kind == GP_EDGE_VR && has_b && (s.gamma[p] == 0.0 || s.par_row[p] < 0 || s.out_row[p] == s.edge_row[p])) ||
                has_b != has_e || s.bus_row[p] >= d->n_vrows || s.par_row[p] >= d->n_vrows ||
                s.edge_row[p] >= nx || s.load_srow[p] >= d->n_srows || s.bus_row[p] < -1 ||
                s.par_row[p] < -1 || s.load_srow[p] < -1) {
                delete im;
                return fail(GP_EINVAL, "gp_nr3_create: step %d: a bus must carry exactly the phases of "
                                       "its incoming edge, with valid row indices", i);
            }
            s.cap[p] = d->cappu[3 * i + p];
            if (has_b) {
                vrow_srow[s.bus_row[p]] = s.load_srow[p];
                vrow_cap[s.bus_row[p]] = s.cap[p];
            }
            for (int q = 0; q < 3; ++q) {
                s.z[3 * p + q].x = d->z[((size_t)i * 9 + 3 * p + q) * 2];
                s.z[3 * p + q].y = d->z[((size_t)i * 9 + 3 * p + q) * 2 + 1];
            }
        }
    }
    for (Nr3Step& st : steps)      // rows of a per-scenario gamma array: step, then phase order
        for (int p = 0; p < 3; ++p)
            st.gam_row[p] = (st.kind == GP_EDGE_VR && st.gamma[p] != 0.0) ? im->n_grows++ : -1;
    {
        int n3 = 0;
        for (const Nr3Step& st : steps) n3 += st.bus_row[0] >= 0 && st.bus_row[1] >= 0 && st.bus_row[2] >= 0;
        im->mostly_3ph = 6 * n3 >= 5 * d->n_steps;
    }
    std::vector<int4> childs((size_t)d->n_child);
    for (int c = 0; c < d->n_child; ++c) {
        childs[c].x = d->child_step[c];
        int* rows[3] = {&childs[c].y, &childs[c].z, &childs[c].w};
        for (int p = 0; p < 3; ++p) {
            const int r = d->child_irow[3 * c + p];
            *rows[p] = r >= 0 ? r + d->n_vrows : -1;
        }
        if (childs[c].x < 0 || childs[c].x >= d->n_steps) {
            delete im;
            return fail(GP_EINVAL, "gp_nr3_create: child %d step index out of range", c);
        }
    }
    for (int p = 0; p < 3; ++p) {
        im->k.root_row[p] = d->root_vrow[p];
        if (d->root_vrow[p] < 0 || d->root_vrow[p] >= d->n_vrows) {
            delete im;
            return fail(GP_EINVAL, "gp_nr3_create: the slack bus must carry all three phases");
        }
        im->k.vslack[p] = make_double2(d->vslack[2 * p], d->vslack[2 * p + 1]);
        im->k.cA[p] = d->cA[p];
        im->k.cB[p] = d->cB[p];
        vrow_srow[d->root_vrow[p]] = d->root_load_srow[p];
        vrow_cap[d->root_vrow[p]] = d->root_cappu[p];
    }
    im->k.n_steps = d->n_steps;
    im->k.aPQ = d->zip[0];
    im->k.aI = d->zip[1];
    im->k.aZ = d->zip[2];
    im->k.betaS = d->zip[0];
    im->k.f0_absent = d->f0_absent;
    std::vector<int> lines((size_t)d->n_lines * 9);
    for (int l = 0; l < d->n_lines; ++l)
        for (int p = 0; p < 3; ++p) {
            const int ir = d->line_irow[3 * l + p];
            lines[9 * l + p] = ir >= 0 ? ir + d->n_vrows : -1;
            lines[9 * l + 3 + p] = d->line_tx_vrow[3 * l + p];
            lines[9 * l + 6 + p] = d->line_rx_vrow[3 * l + p];
        }
    auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, bytes ? bytes : 16);
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t e = up((void**)&im->d_steps, steps.data(), steps.size() * sizeof(Nr3Step));
    if (e == cudaSuccess) e = up((void**)&im->d_child, childs.data(), childs.size() * sizeof(int4));
    if (e == cudaSuccess) e = up((void**)&im->d_lines, lines.data(), lines.size() * sizeof(int));
    if (e == cudaSuccess) e = up((void**)&im->d_vrow_srow, vrow_srow.data(), vrow_srow.size() * sizeof(int));
    if (e == cudaSuccess) e = up((void**)&im->d_vrow_cap, vrow_cap.data(), vrow_cap.size() * sizeof(double));
    if (e != cudaSuccess) {
        delete im;
        return fail(GP_ECUDA, "gp_nr3_create: %s", cudaGetErrorString(e));
    }
    GpFeeder* f = new (std::nothrow) GpFeeder();
    f->device = device;
    f->kind = 2;
    f->impl = im;
    *out = f;
    return GP_OK;
}

void gp_nr3_destroy_impl(void* impl) {
    Nr3Impl* im = (Nr3Impl*)impl;
    cudaFree(im->d_steps);
    cudaFree(im->d_child);
    cudaFree(im->d_lines);
    cudaFree(im->d_vrow_srow);
    cudaFree(im->d_vrow_cap);
    delete im;
}

// gp_set_option("nr3_prefetch", d): -1 = automatic (on for batches that leave the SMs latency-bound)
namespace gp {
int g_nr3_prefetch = -1;
int g_nr3_ring = -1;     // gp_set_option("nr3_ring", -1|0|1): the phase-lane kernel stages next-step rows in shared memory
int g_nr3_kernel = 0;
int nr3_gamma_rows(GpFeeder* f) { return ((Nr3Impl*)f->impl)->n_grows; }
}
static int gp_nr3_prefetch_distance(int64_t S) {
    if (g_nr3_prefetch >= 0) return g_nr3_prefetch;
    return S <= 65536 ? 2 : 0;
}

static int nr3_check(GpFeeder* f, int64_t S, int64_t ld, const char* who) {
    if (!f || f->kind != 2) return fail(GP_EINVAL, "%s: not an NR3 feeder handle", who);
    if (S < 0 || ld < S) return fail(GP_EINVAL, "%s: need 0 <= S <= ld (S=%lld ld=%lld)", who, (long long)S, (long long)ld);
    return GP_OK;
}

extern "C" int64_t gp_nr3_workspace_bytes(GpFeeder* f, int64_t ld) {
    if (!f || f->kind != 2 || ld < 0) return fail(GP_EINVAL, "gp_nr3_workspace_bytes: bad argument");
    Nr3Impl* im = (Nr3Impl*)f->impl;
    return ((int64_t)im->n_fronts * FRONT_D2 + im->n_vrows) * ld * 16 + 256;
}

extern "C" int gp_nr3_solve_batch(GpFeeder* f, int64_t S, int64_t ld, const double* spu, double* X,
                                  int32_t* iters, int32_t* status, double* fmax, void* workspace,
                                  int64_t workspace_bytes, int32_t maxiter, void* stream) {
    return gp_nr3_solve_batch_ex(f, S, ld, spu, nullptr, X, iters, status, fmax, workspace, workspace_bytes, maxiter,
                                 stream);
}

extern "C" int gp_nr3_solve_batch_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const GpSolveExtras* ex,
                                     double* X, int32_t* iters, int32_t* status, double* fmax, void* workspace,
                                     int64_t workspace_bytes, int32_t maxiter, void* stream) {
    int rc = nr3_check(f, S, ld, "gp_nr3_solve_batch");
    if (rc) return rc;
    if (S == 0) return GP_OK;
    Nr3Impl* im = (Nr3Impl*)f->impl;
    const double* gam = (ex && im->n_grows > 0) ? ex->gamma_dev : nullptr;
    const int warm = ex ? ex->warm_start : 0;
    if (warm != 0 && warm != 1) return fail(GP_EINVAL, "gp_nr3_solve_batch_ex: warm_start must be 0 or 1");
    const double* wpu = ex ? ex->wpu_dev : nullptr;
    const bool extras = gam || warm || wpu;
    if (!X || !iters || !status || !fmax || !workspace || (!spu && im->n_srows > 0))
        return fail(GP_EINVAL, "gp_nr3_solve_batch: null buffer");
    if (workspace_bytes < gp_nr3_workspace_bytes(f, ld))
        return fail(GP_EINVAL, "gp_nr3_solve_batch: workspace too small (%lld < %lld bytes)",
                    (long long)workspace_bytes, (long long)gp_nr3_workspace_bytes(f, ld));
    GP_CUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int BLOCK = 64;
    const unsigned gx = (unsigned)((S + BLOCK - 1) / BLOCK);
    double2* W = (double2*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double2* DV = W + (size_t)im->n_fronts * FRONT_D2 * ld;
    // nr3_kernel option: 0 automatic (phase-parallel lanes on Line / Transformer feeders, one thread
    // per scenario when the feeder has regulator edges), 1 thread per scenario, 2 phase-parallel
    // Measured on B200 (round 1b): the phase-parallel kernel wins where a thread per scenario leaves
    // the SMs nearly empty - C3 (37-bus, 16,384 scenarios) 1.15 ms vs 1.48 ms, 123-bus 17,520 scenarios
    // 3.38 ms vs 3.72 ms; once a thread per scenario fills the machine (8 warps x 148 SMs = 37,888
    // scenarios) it is 25-35 % faster (fewer shuffles, full 512-byte row pieces).
    const bool phase = !extras && (g_nr3_kernel == 2 || (g_nr3_kernel == 0 && im->lines_only && S <= 32768));
    if (extras) {
        nr3_solve_kernel<BLOCK, true><<<gx, BLOCK, 0, st>>>(im->d_steps, im->d_child, im->k, im->n_vrows, im->n_irows, S,
                                                            ld, (const double2*)spu, (double2*)X, W, DV, iters, status,
                                                            fmax, maxiter, gp_nr3_prefetch_distance(S), gam, warm, wpu);
    } else if (phase) {
        if (!im->lines_only)
            return fail(GP_EINVAL, "gp_nr3_solve_batch: the phase-parallel kernel does not handle regulator edges");
        constexpr int PB = 128;
        const long long warps = (S + 9) / 10;
        const unsigned gp2 = (unsigned)((warps * 32 + PB - 1) / PB);
        // cp.async ring ("nr3_ring": -1 automatic, 0, 1).  Measured on B200, round 2 (profiles/r02j_*): C3 (37-bus, every
        // bus three-phase) 1.115 -> 1.061 ms; C4-NR3 (123-bus, 55 of 125 buses single-phase, whose 2x2 steps are too
        // short to pay for staging ten rows) 3.21 -> 3.35 ms: automatic = on when at least 5 of 6 buses are three-phase
        if (g_nr3_ring == 1 || (g_nr3_ring < 0 && im->mostly_3ph)) {
            constexpr size_t smem = (size_t)(PB / 32) * 3 * 10 * 32 * sizeof(double2);     // one 15 kB ring per warp
            static bool attr_set = false;
            if (!attr_set) {
                GP_CUDA(cudaFuncSetAttribute(nr3_phase_kernel<PB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_set = true;
            }
            nr3_phase_kernel<PB, true><<<gp2, PB, smem, st>>>(im->d_steps, im->d_child, im->k, S, ld, (const double2*)spu,
                                                              (double2*)X, W, DV, iters, status, fmax, maxiter, 0);
        } else {
            nr3_phase_kernel<PB, false><<<gp2, PB, 0, st>>>(im->d_steps, im->d_child, im->k, S, ld, (const double2*)spu,
                                                            (double2*)X, W, DV, iters, status, fmax, maxiter,
                                                            gp_nr3_prefetch_distance(S));
        }
    } else {
        nr3_solve_kernel<BLOCK, false><<<gx, BLOCK, 0, st>>>(im->d_steps, im->d_child, im->k, im->n_vrows, im->n_irows, S,
                                                             ld, (const double2*)spu, (double2*)X, W, DV, iters, status,
                                                             fmax, maxiter, gp_nr3_prefetch_distance(S), nullptr, 0, nullptr);
    }
    count_launch();
    GP_CUDA(cudaGetLastError());
    return GP_OK;
}

extern "C" int gp_nr3_map_output(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const double* X,
                                 double* V, double* I, double* Stx, double* Srx, double* sV, double* iN,
                                 double* Vmag, double* loss, void* stream) {
    return gp_nr3_map_output_ex(f, S, ld, spu, nullptr, X, V, I, Stx, Srx, sV, iN, Vmag, loss, stream);
}

extern "C" int gp_nr3_map_output_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu, const GpSolveExtras* ex,
                                    const double* X, double* V, double* I, double* Stx, double* Srx, double* sV,
                                    double* iN, double* Vmag, double* loss, void* stream) {
    const double* wpu = ex ? ex->wpu_dev : nullptr;
    int rc = nr3_check(f, S, ld, "gp_nr3_map_output");
    if (rc) return rc;
    if (S == 0) return GP_OK;
    Nr3Impl* im = (Nr3Impl*)f->impl;
    if (!X) return fail(GP_EINVAL, "gp_nr3_map_output: null state buffer");
    if ((sV || iN) && !spu && im->n_srows > 0) return fail(GP_EINVAL, "gp_nr3_map_output: sV needs spu");
    GP_CUDA(cudaSetDevice(f->device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int BLOCK = 128;
    const unsigned gx = (unsigned)((S + BLOCK - 1) / BLOCK);
    if (V || I || sV || iN || Vmag) {
        dim3 g(gx, 8);
        nr3_map_rows_kernel<<<g, BLOCK, 0, st>>>(im->n_vrows, im->n_irows, im->d_vrow_srow, im->d_vrow_cap, im->k,
                                                 S, ld, (const double2*)spu, (const double2*)X, (double2*)V,
                                                 (double2*)I, (double2*)sV, (double2*)iN, Vmag, wpu);
        count_launch();
    }
    if (Stx || Srx || loss) {
        nr3_map_lines_kernel<<<gx, BLOCK, 0, st>>>(im->n_lines, im->n_vrows, im->d_lines, S, ld,
                                                   (const double2*)X, (double2*)Stx, (double2*)Srx, (double2*)loss);
        count_launch();
    }
    GP_CUDA(cudaGetLastError());
    return GP_OK;
}
