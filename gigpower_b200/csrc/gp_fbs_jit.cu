// Feeder-specialised on-chip FBS kernel: source generator and NVRTC compilation.
//
// The generic kernels in gp_fbs.cu interpret the feeder's step table: per tree step they fetch
// row indices, phase masks and the 3x3 impedance block, branch on the edge kind and compute
// addresses.  For a feeder small enough to run out of shared memory that interpretation is
// more than half of the instruction stream and sits on the critical path of a solve (the
// kernel has only a handful of warps per SM to hide it with).  The sweep schedule of a feeder is
// fixed once it is flattened, so this file writes it out as straight-line CUDA C++: every row
// offset is an immediate, every impedance a literal, absent phases and load-free bus-phases
// emit no code at all, and the loads of a scenario stay in registers for the whole solve.
// The arithmetic per step is the one of fbs_fwd_step / fbs_bwd_step (same operations, same
// order as the mask-specialised bodies), i.e. solution_fbs.py:87-128 of the reference.
//
// Scheduling, memory layout and outputs are those of fbs_onchip_kernel: one persistent block
// per SM, a 32-scenario shared-memory slice per warp ([row][lane], 512-byte pitch), tiles
// pulled from a device counter, one coalesced write of the final V and I rows.
//
// The source is compiled for sm_100a with NVRTC (loaded with dlopen, so the library has no
// link-time dependency on it) into a cubin that gp_fbs_specialise loads through the driver API.
#include "gp_fbs.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <functional>

namespace gp {

namespace {

struct Src {
    std::string s;
    bool overflow = false;
    void operator()(const char* fmt, ...) {
        char buf[8192];
        va_list ap;
        va_start(ap, fmt);
        const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        if (n >= (int)sizeof(buf)) overflow = true;
        s += buf;
    }
};

// exact literal of a double (hexadecimal floating constant, C++17)
std::string lit(double x) {
    char buf[64];
    if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
    snprintf(buf, sizeof(buf), "%a", x);
    return std::string("(") + buf + ")";
}

}  // namespace

// Shared-memory slots of the I rows.  Two kinds of row need no storage of their own:
//  * a bus-phase without load and capacitor whose current comes from exactly one child edge carries
//    that child's current unchanged (I = 0 + I_child, update_current :256-288): the row shares the
//    child's slot and its store is dropped (the value is already there);
//  * without load, capacitor and any child current the row stays at the flat start's zero.
// IEEE 13-bus: 35 rows -> 24 slots, which is what lets a seventh warp fit into an SM.
// slot[r] = shared-memory slot of I row r, -1 = constant zero; owner[r] = the row's step stores it.
int fbs_jit_islots(const std::vector<FbsStep>& steps, const std::vector<int4>& child, int n_irows,
                   std::vector<int>& slot, std::vector<char>& owner) {
    slot.assign((size_t)(n_irows > 0 ? n_irows : 0), -2);
    owner.assign(slot.size(), 0);
    int n_slots = 0;
    for (int i = (int)steps.size() - 1; i >= 0; --i) {      // children before parents
        const FbsStep& st = steps[i];
        if ((st.kind & 0xff) == GP_EDGE_VR) continue;
        for (int p = 0; p < 3; ++p) {
            const int r = st.edge_irow[p];
            if (r < 0 || r >= n_irows) continue;
            const bool has_load = st.bus_vrow[p] >= 0 && st.load_srow[p] >= 0;
            const bool has_cap = st.bus_vrow[p] >= 0 && st.cap[p] != 0.0;
            int n_src = 0, src = -1;
            bool known = true;
            for (int c = 0; c < st.child_count; ++c) {
                const int4 cr = child[st.child_begin + c];
                const int crow = p == 0 ? cr.x : (p == 1 ? cr.y : cr.z);
                if (crow < 0) continue;
                if (crow >= n_irows || slot[crow] == -2) known = false;     // not a row of a later step: keep apart
                else if (slot[crow] >= 0) { ++n_src; src = slot[crow]; }
            }
            if (!known || has_load || has_cap || n_src > 1) { slot[r] = n_slots++; owner[r] = 1; }
            else if (n_src == 1) slot[r] = src;
            else slot[r] = -1;
        }
    }
    for (size_t r = 0; r < slot.size(); ++r)
        if (slot[r] == -2) { slot[r] = n_slots++; owner[r] = 1; }           // rows no step owns (defensive)
    return n_slots;
}

namespace {
// Tree of the step table for the depth-first build: parent step (-1 = root) and children latest-in-topo-order first.
// false when the table is not a tree in topo order or a regulator leaves a phase of its bus unregulated.
bool dfs_tree(const std::vector<FbsStep>& steps, const FbsConst& k, int n_vrows, std::vector<int>& par,
              std::vector<std::vector<int>>& kids, std::vector<int>& size) {
    const int n = (int)steps.size();
    std::vector<int> row_step((size_t)n_vrows, -2);
    for (int p = 0; p < 3; ++p) row_step[k.root_vrow[p]] = -1;
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < 3; ++p)
            if (steps[i].bus_vrow[p] >= 0) row_step[steps[i].bus_vrow[p]] = i;
    par.assign((size_t)n, -2);
    kids.assign((size_t)n + 1, {});
    for (int i = 0; i < n; ++i) {
        for (int p = 0; p < 3; ++p)
            if (steps[i].par_vrow[p] >= 0) par[i] = row_step[steps[i].par_vrow[p]];
        if (par[i] < -1 || par[i] >= i) return false;
        kids[par[i] + 1].push_back(i);
        if ((steps[i].kind & 0xff) == GP_EDGE_VR)
            for (int p = 0; p < 3; ++p)
                if (steps[i].bus_vrow[p] >= 0 && steps[i].gamma[p] == 0.0) return false;
    }
    for (auto& ks : kids) std::sort(ks.begin(), ks.end(), std::greater<int>());
    size.assign((size_t)n, 1);
    for (int i = n - 1; i >= 0; --i)
        if (par[i] >= 0) size[par[i]] += size[i];
    return true;
}

struct DfsLive { std::vector<std::pair<int, int>> vals; };     // (kind 0 = f / 1 = a, phase) of the parent kept across a child

// What of bus k (step index, -1 = root) must survive the visit of its j-th child (see fbs_jit_source, DFS build)
DfsLive dfs_live(const std::vector<FbsStep>& steps, const std::vector<int>& islot, const std::vector<std::vector<int>>& kids,
                 const std::vector<std::array<bool, 3>>& win, int k, int j) {
    DfsLive L;
    const auto& ks = kids[k + 1];
    for (int p = 0; p < 3; ++p) {
        if (k >= 0 && steps[k].bus_vrow[p] < 0) continue;
        bool needed_later = false, rewritten = false;
        for (int j2 = 0; j2 < (int)ks.size(); ++j2) {
            const FbsStep& c2 = steps[ks[j2]];
            if (j2 > j && c2.par_vrow[p] >= 0 && (c2.bus_vrow[p] >= 0 || (c2.kind & 0xff) == GP_EDGE_VR)) needed_later = true;
            if (j2 >= j && win[ks[j2]][p]) rewritten = true;
        }
        if (needed_later || !rewritten) L.vals.push_back({0, p});
    }
    if (k >= 0 && (steps[k].kind & 0xff) != GP_EDGE_VR)
        for (int p = 0; p < 3; ++p) {
            const int r = steps[k].edge_irow[p];
            if (r < 0 || islot[r] < 0) continue;
            bool earlier = false;
            for (int j2 = 0; j2 < j; ++j2) {
                const FbsStep& c2 = steps[ks[j2]];
                if ((c2.kind & 0xff) != GP_EDGE_VR && c2.edge_irow[p] >= 0 && islot[c2.edge_irow[p]] >= 0) earlier = true;
            }
            if (earlier) L.vals.push_back({1, p});
        }
    return L;
}
}  // namespace

// Shared-memory stack slots the depth-first build needs (-1: the feeder cannot be built that way)
int fbs_jit_dfs_stack(const std::vector<FbsStep>& steps, const std::vector<int4>& child, const FbsConst& k, int n_vrows,
                      int n_irows) {
    const int n = (int)steps.size();
    if (n == 0 || n > 64) return -1;         // straight-line code of a larger feeder outgrows the instruction cache
    std::vector<int> par, size, islot;
    std::vector<char> iowner;
    std::vector<std::vector<int>> kids;
    if (!dfs_tree(steps, k, n_vrows, par, kids, size)) return -1;
    fbs_jit_islots(steps, child, n_irows, islot, iowner);
    std::vector<std::array<bool, 3>> win((size_t)n);
    {
        std::vector<char> written((size_t)n_vrows, 0);
        for (int i = 0; i < n; ++i)
            for (int p = 0; p < 3; ++p) {
                const bool vr = (steps[i].kind & 0xff) == GP_EDGE_VR;
                const bool stores = steps[i].par_vrow[p] >= 0 && (vr ? steps[i].gamma[p] != 0.0 : steps[i].bus_vrow[p] >= 0);
                win[i][p] = stores && !written[steps[i].par_vrow[p]];
                if (stores) written[steps[i].par_vrow[p]] = 1;
            }
    }
    int nst = 0;
    std::function<void(int, int)> walk = [&](int kk, int sp) {
        const auto& ks = kids[kk + 1];
        for (int j = 0; j < (int)ks.size(); ++j) {
            const int c = ks[j];
            int m = 0;
            if (size[c] >= 2) m = (int)dfs_live(steps, islot, kids, win, kk, j).vals.size();
            nst = std::max(nst, sp + m);
            walk(c, sp + m);
        }
    };
    walk(-1, 0);
    return nst;
}

std::string fbs_jit_source(const std::vector<FbsStep>& steps, const std::vector<int4>& child,
                           const std::vector<int>& lines, const FbsConst& k, int n_vrows, int n_irows, int n_srows,
                           int nw, bool ig, bool read_ahead, bool ex, bool dfs) {
    // dfs: the depth-first build.  The sweep walks the tree depth-first (a bus's forward step on the way down, its
    // backward step on the way up, children latest-in-topo-order first - the order fbs_solve_kernel uses, see there
    // for why every value is the reference's) and keeps a bus's voltage in a C++ variable of the block that is its
    // subtree, so V needs NO shared memory: a row goes to the caller's V array once per sweep when it is final, and
    // the epilogue reads it back from there.  Shared memory holds the I slots and a small stack: what a bus must
    // keep across the visit of a child with children of its own (its forward voltage if later children still read
    // it, its partial current sum) is parked there around that child, so the registers are free inside the subtree.
    // 34-bus feeder: 144 -> 66 rows per warp, 3 -> 6 warps per SM; 37-bus: 194 -> 95 rows, 2 -> 4 warps.
    if (dfs && (ig || ex || read_ahead)) return std::string();
    const int n_stack = dfs ? fbs_jit_dfs_stack(steps, child, k, n_vrows, n_irows) : 0;
    if (dfs && n_stack < 0) return std::string();
    // ig: the I rows live in global memory (the caller's I array, L2-resident) instead of shared
    // memory.  Only the V rows then take shared memory, so about twice as many warps fit per SM.
    // The forward sweep's I loads have static addresses and are issued two steps ahead; the
    // backward sweep hands a child's new current to its parent in registers and only stores it.
    Src o;
    const int n_steps = (int)steps.size();
    for (const FbsStep& st : steps)
        for (int j = 0; j < 9; ++j)
            if (!std::isfinite(st.z[j].x) || !std::isfinite(st.z[j].y)) return std::string();
    const bool spu_regs = n_srows <= 32;      // <= 128 registers of loads held for the whole solve
    o("// generated by libgigpower_b200 (gp_fbs_jit.cu): %d steps, %d V rows, %d I rows, %d spu rows\n", n_steps,
      n_vrows, n_irows, n_srows);
    // I row -> storage: shared-memory slots with aliasing (fbs_jit_islots) or, with ig, the caller's rows as they are
    std::vector<int> islot;
    std::vector<char> iowner;
    int n_islots = n_irows;
    if (ig) {
        islot.resize((size_t)n_irows);
        for (int r = 0; r < n_irows; ++r) islot[r] = r;
        iowner.assign((size_t)n_irows, 1);
    } else {
        n_islots = fbs_jit_islots(steps, child, n_irows, islot, iowner);
    }
    auto ild = [&](int r) {                 // expression for the value of I row r
        if (islot[r] < 0) return std::string("zero");
        char b[48];
        snprintf(b, sizeof(b), "ILD(%d)", islot[r]);
        return std::string(b);
    };
    o("#define NW %d\n#define NV %d\n#define NI %d\n#define NIP %d\n#define NS %d\n", nw, n_vrows, n_irows, n_islots,
      n_srows);
    o("#define IG_MODE %d\n", ig ? 1 : 0);
    o("#define DFS_MODE %d\n#define NST %d\n#define NVS %s\n", dfs ? 1 : 0, n_stack, dfs ? "0" : "NV");
    // EX: the build behind gp_fbs_solve_batch_ex - per-scenario regulator ratios (GpSolveExtras.gamma_dev: the sweep
    // reads this scenario's ratio where the plain build has a constant; 1 / gamma is formed first, as 1/gamma * V is in
    // Python, solution_fbs.py:160, 205) and warm starts (the V rows and I slots are loaded from the caller's arrays
    // instead of the flat start, also when a non-finite current sends the scenario to the exact path)
    o("#define EX_MODE %d\n", ex ? 1 : 0);
    o("#define GP_SWEEP_EXTRA_PARAM %s\n#define GP_SWEEP_EXTRA_ARG %s\n",
      ex ? ", const char* __restrict__ gp" : (dfs ? ", char* __restrict__ Vw" : ""), ex ? ", gp" : (dfs ? ", Vw" : ""));
    o("#define READ_AHEAD %d\n", (read_ahead && n_srows <= 32) ? 1 : 0);   // the next tile's loads are held in registers
    o("struct GpPeerDst { double* vmag[%d]; double2* loss[%d]; int n; int pad; };\n", GP_MAX_PEERS, GP_MAX_PEERS);
    if (ig)
        o("#define IPTR char*\n#define SROWS NV\n"
          "#define ILD(r) __ldcg(reinterpret_cast<const double2*>(Ip + (size_t)(r) * (size_t)ldb))\n"
          "#define IST(r, v) (*reinterpret_cast<double2*>(Ip + (size_t)(r) * (size_t)ldb) = (v))\n");
    else
        o("#define IPTR double2*\n#define SROWS (NVS + NIP + NST)\n"
          "#define ILD(r) Ip[(r) * 32]\n"
          "#define IST(r, v) (Ip[(r) * 32] = (v))\n"
          "#define SLD(j) Ip[(NIP + (j)) * 32]\n"
          "#define SST(j, v) (Ip[(NIP + (j)) * 32] = (v))\n");
    if (dfs)
        o("#define VLD(r) __ldcg(reinterpret_cast<const double2*>(Vw + (size_t)(r) * (size_t)ldb))\n"
          "#define VST(r, v) (*reinterpret_cast<double2*>(Vw + (size_t)(r) * (size_t)ldb) = (v))\n");
    else
        o("#define VLD(r) Vs[(r) * 32]\n");
    std::vector<int> row2step((size_t)(n_irows > 0 ? n_irows : 1), -1);
    for (int i = 0; i < n_steps; ++i)
        for (int p = 0; p < 3; ++p)
            if (steps[i].edge_irow[p] >= 0 && steps[i].edge_irow[p] < n_irows) row2step[steps[i].edge_irow[p]] = i;
    o("__device__ __forceinline__ double gp_n2n(double x) {\n"
      "    if (x != x) return 0.0;\n"
      "    if (isinf(x)) return copysign(1.7976931348623157e308, x);\n"
      "    return x;\n}\n");
    // |V| and 1/|V|^2 from one reciprocal-square-root seed, branch-free (the library sqrt and
    // division each carry a slow-path branch, which would cut the sweep into small basic blocks
    // and stop the compiler from overlapping independent subtrees).  Two Newton steps take the
    // 2^-22 seed to full precision; one residual correction each for the root and the reciprocal.
    o("__device__ __forceinline__ void gp_abs_inv(double m2, double& a, double& inv) {\n"
      "    double y;\n"
      "    asm(\"rsqrt.approx.ftz.f64 %%0, %%1;\" : \"=d\"(y) : \"d\"(m2));\n"
      "    y = fma(0.5 * y, fma(-(m2 * y), y, 1.0), y);\n"
      "    y = fma(0.5 * y, fma(-(m2 * y), y, 1.0), y);\n"
      "    double r = m2 * y;\n"
      "    r = fma(0.5 * y, fma(-r, r, m2), r);\n"
      "    double i = y * y;\n"
      "    i = fma(i, fma(-m2, i, 1.0), i);\n"
      "    const bool pos = m2 > 0.0;\n"
      "    a = pos ? r : 0.0;\n"
      "    inv = pos ? i : 0.0;\n"
      "}\n");
    // (L2-coherent loads: in the one-launch host path the rows were written by this kernel)
    o("#define LDS_(r) __ldcg(reinterpret_cast<const double2*>(sp + (size_t)(r) * (size_t)ldb))\n");
    // Impedances, capacitors and regulator ratios go to the constant bank: an FP64 instruction
    // takes a constant-bank operand directly, a 64-bit literal would cost two moves per use.
    o("__constant__ double ZR[%d] = {", n_steps * 9 + 1);
    for (int i = 0; i < n_steps; ++i)
        for (int j = 0; j < 9; ++j) o("%s,", lit(steps[i].z[j].x).c_str());
    o("0.0};\n__constant__ double ZI[%d] = {", n_steps * 9 + 1);
    for (int i = 0; i < n_steps; ++i)
        for (int j = 0; j < 9; ++j) o("%s,", lit(steps[i].z[j].y).c_str());
    o("0.0};\n__constant__ double CAP[%d] = {", n_steps * 3 + 1);
    for (int i = 0; i < n_steps; ++i)
        for (int j = 0; j < 3; ++j) o("%s,", lit(steps[i].cap[j]).c_str());
    o("0.0};\n__constant__ double GAM[%d] = {", n_steps * 6 + 1);
    for (int i = 0; i < n_steps; ++i) {
        for (int j = 0; j < 3; ++j) o("%s,", lit(steps[i].gamma[j]).c_str());
        for (int j = 0; j < 3; ++j) o("%s,", lit(steps[i].inv_gamma[j]).c_str());
    }
    o("0.0};\n__constant__ double KC[9] = {%s,%s,%s,%s,%s,%s,%s,%s,%s};\n", lit(k.aPQ).c_str(), lit(k.aI).c_str(),
      lit(k.aZ).c_str(), lit(k.vref[0].x).c_str(), lit(k.vref[0].y).c_str(), lit(k.vref[1].x).c_str(),
      lit(k.vref[1].y).c_str(), lit(k.vref[2].x).c_str(), lit(k.vref[2].y).c_str());
    // One full iteration (forward sweep, backward sweep, root mismatch) as a function that is NOT
    // inlined: inside the solve loop the compiler would hoist the ~300 loop-invariant constant
    // loads out of the loop and spill them; as a function body they are fetched where they are used.
    // EXACT = false: branch-free arithmetic; a non-finite current raises `bad` instead of being
    // repaired, and the caller restarts that scenario with EXACT = true, whose sqrt, division,
    // V != 0 test and nan_to_num are literally those of fbs_bwd_step.  For every trajectory that
    // stays finite the two agree to rounding, so the restart only costs time on diverging scenarios.
    o("template <bool EXACT>\n"
      "__device__ __noinline__ double gp_sweep(double2* __restrict__ Vs, IPTR __restrict__ Ip,\n"
      "                                        const char* __restrict__ sp, unsigned ldb, bool& bad_out GP_SWEEP_EXTRA_PARAM) {\n"
      "    bool bad = false;\n"
      "    const double2 zero = make_double2(0.0, 0.0);\n"
      "    const double aPQ = KC[0], aI = KC[1], aZ = KC[2];\n");
    auto fwd_loads = [&](int i) {       // the I rows forward step i reads, as named values
        if (i < 0 || i >= n_steps || (steps[i].kind & 0xff) == GP_EDGE_VR) return;
        for (int q = 0; q < 3; ++q)
            if (steps[i].edge_irow[q] >= 0) o("    fi%d_%d = ILD(%d);\n", i, q, steps[i].edge_irow[q]);
    };
    if (ig) {
        for (int i = 0; i < n_steps; ++i)
            for (int q = 0; q < 3; ++q)
                if (steps[i].edge_irow[q] >= 0 && (steps[i].kind & 0xff) != GP_EDGE_VR)
                    o("    double2 fi%d_%d, cn%d_%d;\n", i, q, i, q);
        fwd_loads(0);
        fwd_loads(1);
    }
    if (dfs) {
        if (spu_regs)
            for (int r = 0; r < n_srows; ++r) o("    const double2 L%d = LDS_(%d);\n", r, r);
        std::vector<int> par, size;
        std::vector<std::vector<int>> kids;
        dfs_tree(steps, k, n_vrows, par, kids, size);
        std::vector<std::array<bool, 3>> win((size_t)n_steps);
        {
            std::vector<char> written((size_t)n_vrows, 0);
            for (int i = 0; i < n_steps; ++i)
                for (int p = 0; p < 3; ++p) {
                    const bool vr = (steps[i].kind & 0xff) == GP_EDGE_VR;
                    const bool stores = steps[i].par_vrow[p] >= 0 && (vr ? steps[i].gamma[p] != 0.0 : steps[i].bus_vrow[p] >= 0);
                    win[i][p] = stores && !written[steps[i].par_vrow[p]];
                    if (stores) written[steps[i].par_vrow[p]] = 1;
                }
        }
        auto vname = [&](int node, int p) {
            char b[32];
            if (node < 0) snprintf(b, sizeof(b), "fR%d", p);
            else snprintf(b, sizeof(b), "f%d_%d", node, p);
            return std::string(b);
        };
        for (int p = 0; p < 3; ++p) o("    double2 fR%d = make_double2(KC[%d], KC[%d]);\n", p, 3 + 2 * p, 4 + 2 * p);
        // `restore`: what the PARENT parked around this bus's visit; it is taken back after this bus's own children,
        // right before the backward step that adds to the parent's current sum and may rewrite its voltage
        std::function<void(int, int, const std::string&)> emit = [&](int i, int sp, const std::string& restore) {
            const FbsStep& st = steps[i];
            const bool vr = (st.kind & 0xff) == GP_EDGE_VR;
            const int pk = par[i];
            o("    {   // bus of step %d\n", i);
            // ---- on the way down: forward step (update_voltage_forward :162-184 / vr_forward :138-160) ----
            if (!vr)
                for (int q = 0; q < 3; ++q)
                    if (st.edge_irow[q] >= 0 && islot[st.edge_irow[q]] >= 0)
                        o("      const double2 i%d = %s;\n", q, ild(st.edge_irow[q]).c_str());
            for (int p = 0; p < 3; ++p) {
                if (st.bus_vrow[p] < 0) continue;
                const std::string P = st.par_vrow[p] >= 0 ? vname(pk, p) : std::string("zero");
                if (vr) {
                    o("      double2 f%d_%d = make_double2(GAM[%d] * %s.x, GAM[%d] * %s.y);\n", i, p, 6 * i + p, P.c_str(),
                      6 * i + p, P.c_str());
                    continue;
                }
                o("      double2 f%d_%d = %s;\n", i, p, P.c_str());
                for (int q = 0; q < 3; ++q) {
                    if (st.edge_irow[q] < 0 || islot[st.edge_irow[q]] < 0) continue;
                    const int zi = 9 * i + 3 * p + q;
                    o("      if (EXACT) {\n");
                    o("        f%d_%d.x -= ZR[%d] * i%d.x - ZI[%d] * i%d.y;\n", i, p, zi, q, zi, q);
                    o("        f%d_%d.y -= ZR[%d] * i%d.y + ZI[%d] * i%d.x;\n", i, p, zi, q, zi, q);
                    o("      } else {\n");
                    o("        f%d_%d.x = fma(ZI[%d], i%d.y, fma(-ZR[%d], i%d.x, f%d_%d.x));\n", i, p, zi, q, zi, q, i, p);
                    o("        f%d_%d.y = fma(-ZI[%d], i%d.x, fma(-ZR[%d], i%d.y, f%d_%d.y));\n", i, p, zi, q, zi, q, i, p);
                    o("      }\n");
                }
            }
            if (!vr)
                for (int p = 0; p < 3; ++p)
                    if (st.edge_irow[p] >= 0 && islot[st.edge_irow[p]] >= 0) o("      double2 a%d_%d = zero;\n", i, p);
            // ---- the children, latest in topo order first ----
            const auto& ks = kids[i + 1];
            for (int j = 0; j < (int)ks.size(); ++j) {
                const int c = ks[j];
                DfsLive L;
                if (size[c] >= 2) L = dfs_live(steps, islot, kids, win, i, j);
                std::string back;
                for (size_t m = 0; m < L.vals.size(); ++m) {
                    char b[96];
                    o("      SST(%d, %c%d_%d);\n", sp + (int)m, L.vals[m].first ? 'a' : 'f', i, L.vals[m].second);
                    snprintf(b, sizeof(b), "      %c%d_%d = SLD(%d);\n", L.vals[m].first ? 'a' : 'f', i, L.vals[m].second, sp + (int)m);
                    back += b;
                }
                emit(c, sp + (int)L.vals.size(), back);
            }
            o("%s", restore.c_str());
            // ---- on the way up: backward step ----
            if (vr) {       // vr_current leaves Itx = 0; vr_backward :186-205
                for (int p = 0; p < 3; ++p) {
                    if (!win[i][p]) continue;
                    if (st.bus_vrow[p] >= 0)
                        o("      %s = make_double2(GAM[%d] * f%d_%d.x, GAM[%d] * f%d_%d.y);\n", vname(pk, p).c_str(),
                          6 * i + 3 + p, i, p, 6 * i + 3 + p, i, p);
                    else
                        o("      %s = make_double2(GAM[%d] * 0.0, GAM[%d] * 0.0);\n", vname(pk, p).c_str(), 6 * i + 3 + p,
                          6 * i + 3 + p);
                }
                for (int p = 0; p < 3; ++p)
                    if (st.bus_vrow[p] >= 0) o("      VST(%d, f%d_%d);\n", st.bus_vrow[p], i, p);
                o("    }\n");
                return;
            }
            bool any_edge = false;
            for (int p = 0; p < 3; ++p) {
                if (st.edge_irow[p] < 0) continue;
                any_edge = true;
                if (islot[st.edge_irow[p]] >= 0) o("      double2 n%d = a%d_%d;\n", p, i, p);
                else o("      double2 n%d = zero;\n", p);
                const bool has_load = st.bus_vrow[p] >= 0 && st.load_srow[p] >= 0;
                const bool has_cap = st.bus_vrow[p] >= 0 && st.cap[p] != 0.0;
                if (has_load || has_cap) {      // calc_sV (solution.py:177-220) + update_current (:256-288)
                    o("      {\n"
                      "        const double m2 = fma(f%d_%d.x, f%d_%d.x, f%d_%d.y * f%d_%d.y);\n"
                      "        double a, inv;\n"
                      "        if (EXACT) { a = sqrt(m2); inv = 0.0; } else gp_abs_inv(m2, a, inv);\n"
                      "        const double a2 = a * a;\n", i, p, i, p, i, p, i, p);
                    if (has_load) {
                        if (spu_regs) o("        const double2 l = L%d;\n", st.load_srow[p]);
                        else o("        const double2 l = LDS_(%d);\n", st.load_srow[p]);
                        o("        const double f = aPQ + aI * a + aZ * a2;\n"
                          "        double2 sv = make_double2(l.x * f, l.y * f);\n");
                    } else {
                        o("        double2 sv = zero;\n");
                    }
                    o("        sv.y -= CAP[%d] * a2;\n", 3 * i + p);
                    o("        if (EXACT) {\n"
                      "          if (f%d_%d.x != 0.0 || f%d_%d.y != 0.0) {\n"
                      "            inv = 1.0 / m2;\n"
                      "            n%d.x += (sv.x * f%d_%d.x + sv.y * f%d_%d.y) * inv;\n"
                      "            n%d.y -= (sv.y * f%d_%d.x - sv.x * f%d_%d.y) * inv;\n"
                      "          }\n"
                      "        } else {\n"
                      "          n%d.x += (sv.x * f%d_%d.x + sv.y * f%d_%d.y) * inv;\n"
                      "          n%d.y -= (sv.y * f%d_%d.x - sv.x * f%d_%d.y) * inv;\n"
                      "        }\n"
                      "      }\n", i, p, i, p, p, i, p, i, p, p, i, p, i, p, p, i, p, i, p, p, i, p, i, p);
                }
            }
            if (any_edge) {                     // nan_to_num (:288) acts on non-finite values only
                o("      {\n        const bool nf = !(0.0");
                for (int p = 0; p < 3; ++p)
                    if (st.edge_irow[p] >= 0) o(" + fabs(n%d.x) + fabs(n%d.y)", p, p);
                o(" <= 1.7976931348623157e308);\n"
                  "        if (EXACT) {\n"
                  "          if (nf) {\n");
                for (int p = 0; p < 3; ++p)
                    if (st.edge_irow[p] >= 0) o("            n%d = make_double2(gp_n2n(n%d.x), gp_n2n(n%d.y));\n", p, p, p);
                o("          }\n"
                  "        } else {\n"
                  "          bad |= nf;\n"
                  "        }\n"
                  "      }\n");
                for (int p = 0; p < 3; ++p)
                    if (st.edge_irow[p] >= 0 && iowner[st.edge_irow[p]]) o("      IST(%d, n%d);\n", islot[st.edge_irow[p]], p);
            }
            for (int p = 0; p < 3; ++p) {       // the surviving parent rewrites (update_voltage_backward :207-232)
                if (!win[i][p]) continue;
                const std::string P = vname(pk, p);
                o("      %s = f%d_%d;\n", P.c_str(), i, p);
                for (int q = 0; q < 3; ++q) {
                    if (st.edge_irow[q] < 0 || islot[st.edge_irow[q]] < 0) continue;
                    const int zi = 9 * i + 3 * p + q;
                    o("      if (EXACT) {\n");
                    o("        %s.x += ZR[%d] * n%d.x - ZI[%d] * n%d.y;\n", P.c_str(), zi, q, zi, q);
                    o("        %s.y += ZR[%d] * n%d.y + ZI[%d] * n%d.x;\n", P.c_str(), zi, q, zi, q);
                    o("      } else {\n");
                    o("        %s.x = fma(-ZI[%d], n%d.y, fma(ZR[%d], n%d.x, %s.x));\n", P.c_str(), zi, q, zi, q, P.c_str());
                    o("        %s.y = fma(ZI[%d], n%d.x, fma(ZR[%d], n%d.y, %s.y));\n", P.c_str(), zi, q, zi, q, P.c_str());
                    o("      }\n");
                }
            }
            if (pk >= 0 && (steps[pk].kind & 0xff) != GP_EDGE_VR)       // the parent's current sum (:274-284)
                for (int p = 0; p < 3; ++p) {
                    const int r = st.edge_irow[p], rp = steps[pk].edge_irow[p];
                    if (r < 0 || islot[r] < 0 || rp < 0 || islot[rp] < 0) continue;
                    o("      a%d_%d.x += n%d.x; a%d_%d.y += n%d.y;\n", pk, p, p, pk, p, p);
                }
            for (int p = 0; p < 3; ++p)
                if (st.bus_vrow[p] >= 0) o("      VST(%d, f%d_%d);\n", st.bus_vrow[p], i, p);
            o("    }\n");
        };
        {
            const auto& ks = kids[0];
            for (int j = 0; j < (int)ks.size(); ++j) {
                const int c = ks[j];
                DfsLive L;
                if (size[c] >= 2) L = dfs_live(steps, islot, kids, win, -1, j);
                std::string back;
                for (size_t m = 0; m < L.vals.size(); ++m) {
                    char b[96];
                    o("    SST(%d, fR%d);\n", (int)m, L.vals[m].second);
                    snprintf(b, sizeof(b), "      fR%d = SLD(%d);\n", L.vals[m].second, (int)m);
                    back += b;
                }
                emit(c, (int)L.vals.size(), back);
            }
        }
        for (int p = 0; p < 3; ++p) o("    VST(%d, fR%d);\n", k.root_vrow[p], p);
        o("    double diff = 0.0;\n"
          "    bool nan = false;\n");
        for (int p = 0; p < 3; ++p)
            o("    { const double dx = fR%d.x - KC[%d], dy = fR%d.y - KC[%d];\n"
              "      const double d = sqrt(fma(dx, dx, dy * dy)); nan |= (d != d); diff = fmax(diff, d); }\n",
              p, 3 + 2 * p, p, 4 + 2 * p);
        o("    if (nan) diff = __longlong_as_double(0x7ff8000000000000LL);\n"
          "    bad_out = bad;\n"
          "    return diff;\n"
          "}\n");
    } else {
    if (spu_regs)       // the loads of the scenario: in flight during the forward sweep
        for (int r = 0; r < n_srows; ++r) o("    const double2 L%d = LDS_(%d);\n", r, r);
    for (int p = 0; p < 3; ++p)
        o("            Vs[%d] = make_double2(KC[%d], KC[%d]);\n", k.root_vrow[p] * 32, 3 + 2 * p, 4 + 2 * p);
    // ---------------- forward sweep (solution_fbs.py:92-98) ----------------
    for (int i = 0; i < n_steps; ++i) {
        const FbsStep& st = steps[i];
        const int kind = st.kind & 0xff;
        o("            {   // forward step %d\n", i);
        if (ig) fwd_loads(i + 2);
        if (kind == GP_EDGE_VR) {           // vr_forward :138-160
            for (int p = 0; p < 3; ++p) {
                if (st.gamma[p] == 0.0 || st.bus_vrow[p] < 0) continue;
                if (ex)
                    o("                { const double g = gp ? __ldg(reinterpret_cast<const double*>(gp + (size_t)%d * (size_t)(ldb >> 1))) : GAM[%d];\n",
                      st.gam_row[p], 6 * i + p);
                else
                    o("                { const double g = GAM[%d];\n", 6 * i + p);
                if (st.par_vrow[p] >= 0)
                    o("                  const double2 a = Vs[%d]; Vs[%d] = make_double2(g * a.x, g * a.y); }\n",
                      st.par_vrow[p] * 32, st.bus_vrow[p] * 32);
                else
                    o("                  Vs[%d] = make_double2(g * 0.0, g * 0.0); }\n", st.bus_vrow[p] * 32);
            }
            o("            }\n");
            continue;
        }
        for (int q = 0; q < 3; ++q)
            if (st.edge_irow[q] >= 0) {
                if (ig) o("                const double2 i%d = fi%d_%d;\n", q, i, q);
                else o("                const double2 i%d = %s;\n", q, ild(st.edge_irow[q]).c_str());
            }
        for (int p = 0; p < 3; ++p) {       // update_voltage_forward :162-184
            if (st.bus_vrow[p] < 0) continue;
            if (st.par_vrow[p] >= 0) o("                { double2 v = Vs[%d];\n", st.par_vrow[p] * 32);
            else o("                { double2 v = zero;\n");
            for (int q = 0; q < 3; ++q) {
                if (st.edge_irow[q] < 0 || islot[st.edge_irow[q]] < 0) continue;    // Z * 0: nothing to subtract
                const int zi = 9 * i + 3 * p + q;
                // complex multiply-subtract: four fused multiply-adds on the fast path (the
                // three-operation form of fbs_fwd_step, bit for bit, on the exact path)
                o("                  if (EXACT) {\n");
                o("                    v.x -= ZR[%d] * i%d.x - ZI[%d] * i%d.y;\n", zi, q, zi, q);
                o("                    v.y -= ZR[%d] * i%d.y + ZI[%d] * i%d.x;\n", zi, q, zi, q);
                o("                  } else {\n");
                o("                    v.x = fma(ZI[%d], i%d.y, fma(-ZR[%d], i%d.x, v.x));\n", zi, q, zi, q);
                o("                    v.y = fma(-ZI[%d], i%d.x, fma(-ZR[%d], i%d.y, v.y));\n", zi, q, zi, q);
                o("                  }\n");
            }
            o("                  Vs[%d] = v; }\n", st.bus_vrow[p] * 32);
        }
        o("            }\n");
    }
    // ---------------- backward sweep (:104-121) ----------------
    for (int i = n_steps - 1; i >= 0; --i) {
        const FbsStep& st = steps[i];
        const int kind = st.kind & 0xff;
        o("            {   // backward step %d\n", i);
        if (kind == GP_EDGE_VR) {           // vr_current leaves Itx = 0; vr_backward :186-205
            for (int p = 0; p < 3; ++p) {
                if (st.gamma[p] == 0.0 || st.par_vrow[p] < 0) continue;
                if (ex)
                    o("                { const double g = gp ? 1.0 / __ldg(reinterpret_cast<const double*>(gp + (size_t)%d * (size_t)(ldb >> 1))) : GAM[%d];\n",
                      st.gam_row[p], 6 * i + 3 + p);
                else
                    o("                { const double g = GAM[%d];\n", 6 * i + 3 + p);
                if (st.bus_vrow[p] >= 0)
                    o("                  const double2 a = Vs[%d]; Vs[%d] = make_double2(g * a.x, g * a.y); }\n",
                      st.bus_vrow[p] * 32, st.par_vrow[p] * 32);
                else
                    o("                  Vs[%d] = make_double2(g * 0.0, g * 0.0); }\n", st.par_vrow[p] * 32);
            }
            o("            }\n");
            continue;
        }
        for (int p = 0; p < 3; ++p)
            if (st.bus_vrow[p] >= 0) o("                const double2 b%d = Vs[%d];\n", p, st.bus_vrow[p] * 32);
        bool any_edge = false;
        for (int p = 0; p < 3; ++p) {
            if (st.edge_irow[p] < 0) continue;
            any_edge = true;
            o("                double2 n%d = zero;\n", p);
            for (int c = 0; c < st.child_count; ++c) {          // children's currents (:274-284)
                const int4 cr = child[st.child_begin + c];
                const int crow[3] = {cr.x, cr.y, cr.z};
                if (crow[p] >= 0) {
                    if (ig && row2step[crow[p]] >= 0)
                        o("                n%d.x += cn%d_%d.x; n%d.y += cn%d_%d.y;\n", p, row2step[crow[p]], p, p,
                          row2step[crow[p]], p);
                    else
                        o("                { const double2 c = %s; n%d.x += c.x; n%d.y += c.y; }\n", ild(crow[p]).c_str(), p, p);
                }
            }
            // calc_sV (solution.py:177-220) + update_current (:256-288); a bus-phase without load
            // and capacitor has sV = 0 and contributes nothing
            const bool has_load = st.bus_vrow[p] >= 0 && st.load_srow[p] >= 0;
            const bool has_cap = st.bus_vrow[p] >= 0 && st.cap[p] != 0.0;
            if (has_load || has_cap) {
                o("                {\n"
                  "                    const double m2 = fma(b%d.x, b%d.x, b%d.y * b%d.y);\n"
                  "                    double a, inv;\n"
                  "                    if (EXACT) { a = sqrt(m2); inv = 0.0; } else gp_abs_inv(m2, a, inv);\n"
                  "                    const double a2 = a * a;\n", p, p, p, p);
                if (has_load) {
                    if (spu_regs) o("                    const double2 l = L%d;\n", st.load_srow[p]);
                    else o("                    const double2 l = LDS_(%d);\n", st.load_srow[p]);
                    o("                    const double f = aPQ + aI * a + aZ * a2;\n"
                      "                    double2 sv = make_double2(l.x * f, l.y * f);\n");
                } else {
                    o("                    double2 sv = zero;\n");
                }
                o("                    sv.y -= CAP[%d] * a2;\n", 3 * i + p);
                o("                    if (EXACT) {\n"
                  "                        if (b%d.x != 0.0 || b%d.y != 0.0) {\n"
                  "                            inv = 1.0 / m2;\n"
                  "                            n%d.x += (sv.x * b%d.x + sv.y * b%d.y) * inv;\n"
                  "                            n%d.y -= (sv.y * b%d.x - sv.x * b%d.y) * inv;\n"
                  "                        }\n"
                  "                    } else {\n"
                  "                        n%d.x += (sv.x * b%d.x + sv.y * b%d.y) * inv;\n"
                  "                        n%d.y -= (sv.y * b%d.x - sv.x * b%d.y) * inv;\n"
                  "                    }\n"
                  "                }\n", p, p, p, p, p, p, p, p, p, p, p, p, p, p);
            }
        }
        if (any_edge) {                     // nan_to_num (:288) acts on non-finite values only
            o("                {\n                    const bool nf = !(0.0");
            for (int p = 0; p < 3; ++p)
                if (st.edge_irow[p] >= 0) o(" + fabs(n%d.x) + fabs(n%d.y)", p, p);
            o(" <= 1.7976931348623157e308);\n"
              "                    if (EXACT) {\n"
              "                        if (nf) {\n");
            for (int p = 0; p < 3; ++p)
                if (st.edge_irow[p] >= 0) o("                            n%d = make_double2(gp_n2n(n%d.x), gp_n2n(n%d.y));\n", p, p, p);
            o("                        }\n"
              "                    } else {\n"
              "                        bad |= nf;\n"
              "                    }\n"
              "                }\n");
            for (int p = 0; p < 3; ++p)
                if (st.edge_irow[p] >= 0) {
                    if (iowner[st.edge_irow[p]]) o("                IST(%d, n%d);\n", islot[st.edge_irow[p]], p);
                    if (ig) o("                cn%d_%d = n%d;\n", i, p, p);
                }
        }
        for (int p = 0; p < 3; ++p) {       // update_voltage_backward (:207-232)
            if (st.bus_vrow[p] < 0 || st.par_vrow[p] < 0) continue;
            o("                { double2 v = b%d;\n", p);
            for (int q = 0; q < 3; ++q) {
                if (st.edge_irow[q] < 0 || islot[st.edge_irow[q]] < 0) continue;    // a current that is always zero
                const int zi = 9 * i + 3 * p + q;
                o("                  if (EXACT) {\n");
                o("                    v.x += ZR[%d] * n%d.x - ZI[%d] * n%d.y;\n", zi, q, zi, q);
                o("                    v.y += ZR[%d] * n%d.y + ZI[%d] * n%d.x;\n", zi, q, zi, q);
                o("                  } else {\n");
                o("                    v.x = fma(-ZI[%d], n%d.y, fma(ZR[%d], n%d.x, v.x));\n", zi, q, zi, q);
                o("                    v.y = fma(ZI[%d], n%d.x, fma(ZR[%d], n%d.y, v.y));\n", zi, q, zi, q);
                o("                  }\n");
            }
            o("                  Vs[%d] = v; }\n", st.par_vrow[p] * 32);
        }
        o("            }\n");
    }
    // ---------------- convergence (:123-128) ----------------
    o("            double diff = 0.0;\n"
      "            bool nan = false;\n");
    for (int p = 0; p < 3; ++p)
        o("            { const double2 v = Vs[%d]; const double dx = v.x - KC[%d], dy = v.y - KC[%d];\n"
          "              const double d = sqrt(fma(dx, dx, dy * dy)); nan |= (d != d); diff = fmax(diff, d); }\n",
          k.root_vrow[p] * 32, 3 + 2 * p, 4 + 2 * p);
    o("            if (nan) diff = __longlong_as_double(0x7ff8000000000000LL);\n"
      "            bad_out = bad;\n"
      "            return diff;\n"
      "}\n");
    }
    // every I row of the caller's layout from its slot (shared rows are written once per row that carries them)
    o("#define GP_I_ROWS_OUT(dst)");
    for (int r = 0; r < n_irows; ++r)
        o(" \\\n    *reinterpret_cast<double2*>((dst) + (size_t)%d * (size_t)ldb) = %s;", r, ild(r).c_str());
    o("\n");
    // warm start: every I slot from the row that owns it
    o("#define GP_I_SLOTS_IN(src)");
    for (int r = 0; r < n_irows; ++r)
        if (!ig && iowner[r] && islot[r] >= 0)
            o(" \\\n    IST(%d, *reinterpret_cast<const double2*>((src) + (size_t)%d * (size_t)ldb));", islot[r], r);
    o("\n");
    o("extern \"C\" __global__ void __launch_bounds__(NW * 32, 1)\n"
      "gp_fbs_jit_kernel(const char* __restrict__ spu, char* __restrict__ V, char* __restrict__ I,\n"
      "                  int* __restrict__ iters_out, int* __restrict__ status_out, double* __restrict__ diff_out,\n"
      "                  long long S, unsigned ldb, double tol, int maxiter, unsigned* __restrict__ sched,\n"
      "                  double* __restrict__ Vmag, double2* __restrict__ loss, char* __restrict__ stage,\n"
      "                  char* __restrict__ Iout, const GpPeerDst peers\n"
      "#if EX_MODE\n"
      "                  , const char* __restrict__ gam, int warm\n"
      "#endif\n"
      "                  ) {\n"
      "    extern __shared__ __align__(16) double2 gp_smem[];\n"
      "    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;\n"
      "    double2* const Vs = gp_smem + warp * (SROWS * 32) + lane;\n"
      "    const long long n_tiles = (S + 31) >> 5;\n"
      "    const double2 zero = make_double2(0.0, 0.0);\n"
      "    // One-launch host path (stage != NULL: spu, V, ... are mapped HOST memory).  Reads and writes of a\n"
      "    // tile cross PCIe in opposite directions, but a warp that reads, solves and then writes keeps the\n"
      "    // two directions busy one after the other, and so does the whole grid, wave after wave (measured:\n"
      "    // the step takes the SUM of the upload and download times).  With READ_AHEAD a warp takes its next\n"
      "    // tile and issues that tile's PCIe reads BEFORE it issues the stores of the tile it has just solved:\n"
      "    // the upload of tile t+1 then travels while the download of tile t drains.\n"
      "    double2 nx[READ_AHEAD ? NS : 1];\n"
      "    unsigned tile = 0;\n"
      "    if (lane == 0) tile = atomicAdd(&sched[0], 1u);\n"
      "    tile = __shfl_sync(0xffffffffu, tile, 0);\n"
      "    if (READ_AHEAD && stage && (long long)tile < n_tiles) {\n"
      "        const long long sr = (long long)tile * 32 + lane;\n"
      "        const char* np = spu + (sr < S ? sr : S - 1) * 16;\n"
      "        _Pragma(\"unroll\")\n"
      "        for (int r = 0; r < (READ_AHEAD ? NS : 1); ++r)\n"
      "            asm volatile(\"ld.global.v2.f64 {%%0, %%1}, [%%2];\" : \"=d\"(nx[r].x), \"=d\"(nx[r].y)\n"
      "                         : \"l\"(np + (size_t)r * (size_t)ldb) : \"memory\");\n"
      "    }\n"
      "    while ((long long)tile < n_tiles) {\n"
      "        const long long s_raw = (long long)tile * 32 + lane;\n"
      "        const bool live = s_raw < S;\n"
      "        const long long s = live ? s_raw : S - 1;\n"
      "        const char* sp = spu + s * 16;\n"
      "        if (stage) {        // spu is mapped HOST memory: this tile's columns go once over PCIe into the\n"
      "                            // device-side staging rows, which every sweep then re-reads through L2\n"
      "            char* const sd = stage + s * 16;\n"
      "            if (READ_AHEAD) {\n"
      "                if (live) {\n"
      "                    _Pragma(\"unroll\")\n"
      "                    for (int r = 0; r < (READ_AHEAD ? NS : 1); ++r)\n"
      "                        *reinterpret_cast<double2*>(sd + (size_t)r * (size_t)ldb) = nx[r];\n"
      "                }\n"
      "            } else if (live) {\n"
      "                for (int r = 0; r < NS; ++r)\n"
      "                    *reinterpret_cast<double2*>(sd + (size_t)r * (size_t)ldb) =\n"
      "                        *reinterpret_cast<const double2*>(sp + (size_t)r * (size_t)ldb);\n"
      "            }\n"
      "            __syncwarp();\n"
      "            sp = sd;\n"
      "        }\n"
      "        IPTR const Ip = %s;\n"
      "#if DFS_MODE\n"
      "        char* const Vw = V + s * 16;        // every V row goes here once per sweep, when it is final\n"
      "#endif\n"
      "#if EX_MODE\n"
      "        const char* const gp = gam ? gam + s * 8 : nullptr;\n"
      "        if (warm) {     // the caller's V and I rows hold an earlier solution (solution_fbs.py:76-90)\n"
      "            for (int r = 0; r < NV; ++r) Vs[r * 32] = *reinterpret_cast<const double2*>(V + s * 16 + (size_t)r * (size_t)ldb);\n"
      "            { const char* Iw = I + s * 16; GP_I_SLOTS_IN(Iw) }\n"
      "        } else\n"
      "#endif\n"
      "        {\n"
      "        for (int r = 0; r < NVS; ++r) Vs[r * 32] = zero;          // flat start (solution.py:108-125)\n"
      "        for (int r = 0; r < NIP; ++r) IST(r, zero);\n"
      "        }\n"
      "        int it = 0;\n"
      "        double diff = -1.0;\n"
      "        bool converged = false, exact = false;\n"
      "        while (!converged && it < maxiter) {\n"
      "            bool bad = false;\n"
      "            if (exact) diff = gp_sweep<true>(Vs, Ip, sp, ldb, bad GP_SWEEP_EXTRA_ARG);\n"
      "            else diff = gp_sweep<false>(Vs, Ip, sp, ldb, bad GP_SWEEP_EXTRA_ARG);\n"
      "            if (bad) {      // a non-finite current: redo this scenario from its start, exact arithmetic\n"
      "#if EX_MODE\n"
      "                if (warm) {\n"
      "                    for (int r = 0; r < NV; ++r) Vs[r * 32] = *reinterpret_cast<const double2*>(V + s * 16 + (size_t)r * (size_t)ldb);\n"
      "                    { const char* Iw = I + s * 16; GP_I_SLOTS_IN(Iw) }\n"
      "                } else\n"
      "#endif\n"
      "                {\n"
      "                for (int r = 0; r < NVS; ++r) Vs[r * 32] = zero;\n"
      "                for (int r = 0; r < NIP; ++r) IST(r, zero);\n"
      "                }\n"
      "                it = 0;\n"
      "                exact = true;\n"
      "                continue;\n"
      "            }\n"
      "            ++it;\n"
      "            converged = diff <= tol;\n"
      "        }\n"
      "        __syncwarp();\n"
      "        unsigned next = 0;\n"
      "        if (lane == 0) next = atomicAdd(&sched[0], 1u);\n"
      "        next = __shfl_sync(0xffffffffu, next, 0);\n"
      "        if (READ_AHEAD && stage && (long long)next < n_tiles) {\n"
      "            const long long sr = (long long)next * 32 + lane;\n"
      "            const char* np = spu + (sr < S ? sr : S - 1) * 16;\n"
      "            _Pragma(\"unroll\")\n"
      "            for (int r = 0; r < (READ_AHEAD ? NS : 1); ++r)\n"
      "                asm volatile(\"ld.global.v2.f64 {%%0, %%1}, [%%2];\" : \"=d\"(nx[r].x), \"=d\"(nx[r].y)\n"
      "                             : \"l\"(np + (size_t)r * (size_t)ldb) : \"memory\");\n"
      "        }\n"
      "        if (live) {         // final V and I rows: one coalesced 512-byte store per row and warp\n"
      "            char* Vg = V + s * 16;\n"
      "            if (V && !DFS_MODE) for (int r = 0; r < NV; ++r) *reinterpret_cast<double2*>(Vg + (size_t)r * (size_t)ldb) = Vs[r * 32];\n"
      "            // I rows: already in place when they live in the caller's (device) I array; copied out of shared\n"
      "            // memory otherwise; Iout is an optional second destination (mapped host memory in the one-launch path)\n"
      "            if (!IG_MODE && I) { char* Ig = I + s * 16; GP_I_ROWS_OUT(Ig) }\n"
      "            if (Iout) { char* Io = Iout + s * 16; GP_I_ROWS_OUT(Io) }\n"
      "            iters_out[s] = it;\n"
      "            diff_out[s] = diff;\n"
      "            status_out[s] = converged ? %d : ((diff != diff || isinf(diff)) ? %d : %d);\n",
      ig ? "I + s * 16" : "Vs + NVS * 32", GP_ST_CONVERGED, GP_ST_NONFINITE, GP_ST_MAXITER);
    // derived results of gp_fbs_post while the state is still on chip: |V| per V row
    // (calc_Vmag, solution.py:222-224) and loss = sum over real lines of Stx - Srx (:226-233),
    // same operations in the same order as fbs_post_rows_kernel / fbs_post_lines_kernel
    // With peers.n > 0 the same values also go to this rank's block of every rank's gather buffer (the
    // path's all-gather, fused): posted stores over NVLink that drain while the next tile is solved.
    o("            if (peers.n) {\n"
      "                char* Mg = reinterpret_cast<char*>(Vmag) + s * 8;\n"
      "                for (int r = 0; r < NV; ++r) {\n"
      "                    const double2 v = VLD(r);\n"
      "                    const double a = sqrt(fma(v.x, v.x, v.y * v.y));\n"
      "                    const size_t off = (size_t)s * 8 + (size_t)r * (size_t)(ldb >> 1);\n"
      "                    if (Vmag) *reinterpret_cast<double*>(Mg + (size_t)r * (size_t)(ldb >> 1)) = a;\n"
      "                    _Pragma(\"unroll\")\n"
      "                    for (int q = 0; q < %d; ++q)\n"
      "                        if (q < peers.n) *reinterpret_cast<double*>(reinterpret_cast<char*>(peers.vmag[q]) + off) = a;\n"
      "                }\n"
      "            } else if (Vmag) {\n"
      "                char* Mg = reinterpret_cast<char*>(Vmag) + s * 8;\n"
      "                for (int r = 0; r < NV; ++r) {\n"
      "                    const double2 v = VLD(r);\n"
      "                    *reinterpret_cast<double*>(Mg + (size_t)r * (size_t)(ldb >> 1)) = sqrt(fma(v.x, v.x, v.y * v.y));\n"
      "                }\n"
      "            }\n"
      "            if (loss || peers.n) {\n"
      "                double2 acc = zero;\n", GP_MAX_PEERS);
    for (size_t l = 0; l * 9 < lines.size(); ++l)
        for (int p = 0; p < 3; ++p) {
            const int ir = lines[9 * l + p], tr = lines[9 * l + 3 + p], rr = lines[9 * l + 6 + p];
            if (ir < 0) continue;
            o("                { const double2 i = %s;", ild(ir).c_str());
            if (tr >= 0) o(" const double2 vt = VLD(%d);", tr);
            else o(" const double2 vt = zero;");
            if (rr >= 0) o(" const double2 vr = VLD(%d);", rr);
            else o(" const double2 vr = zero;");
            o("\n                  const double2 st = make_double2(vt.x * i.x + vt.y * i.y, vt.y * i.x - vt.x * i.y);\n"
              "                  const double2 sr = make_double2(vr.x * i.x + vr.y * i.y, vr.y * i.x - vr.x * i.y);\n"
              "                  acc.x += st.x - sr.x; acc.y += st.y - sr.y; }\n");
        }
    o("                if (loss) loss[s] = acc;\n"
      "                _Pragma(\"unroll\")\n"
      "                for (int q = 0; q < %d; ++q)\n"
      "                    if (q < peers.n) peers.loss[q][s] = acc;\n"
      "            }\n"
      "        }\n"
      "        tile = next;\n"
      "    }\n", GP_MAX_PEERS);
    o("    if (lane == 0) {\n"
      "        __threadfence();\n"
      "        const unsigned done = atomicAdd(&sched[1], 1u);\n"
      "        if (done == gridDim.x * NW - 1) {\n"
      "            sched[0] = 0u;\n"
      "            sched[1] = 0u;\n"
      "            __threadfence();\n"
      "        }\n"
      "    }\n"
      "}\n");
    return o.overflow ? std::string() : o.s;
}

// ---- NVRTC through dlopen ------------------------------------------------------------------
namespace {
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*DestroyProgram)(nvrtcProgram*);
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.h) break;
    }
    if (!n.h) return n;
#define GP_SYM(f)                                                 \
    *(void**)(&n.f) = dlsym(n.h, "nvrtc" #f);                     \
    if (!n.f) return n;
    GP_SYM(CreateProgram) GP_SYM(CompileProgram) GP_SYM(GetCUBINSize) GP_SYM(GetCUBIN) GP_SYM(GetProgramLogSize)
    GP_SYM(GetProgramLog) GP_SYM(DestroyProgram)
#undef GP_SYM
    n.ok = true;
    return n;
}
}  // namespace

int jit_compile(const std::string& src, std::vector<char>& cubin, std::string& log) {
    Nvrtc& n = nvrtc();
    if (!n.ok) {
        log = "libnvrtc.so.12 could not be loaded";
        return GP_EJIT;
    }
    nvrtcProgram prog = nullptr;
    if (n.CreateProgram(&prog, src.c_str(), "gp_fbs_jit.cu", 0, nullptr, nullptr) != 0) {
        log = "nvrtcCreateProgram failed";
        return GP_EJIT;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    const int rc = n.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    if (n.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
        log.resize(ls);
        n.GetProgramLog(prog, &log[0]);
    }
    if (rc != 0) {
        n.DestroyProgram(&prog);
        return GP_EJIT;
    }
    size_t cs = 0;
    if (n.GetCUBINSize(prog, &cs) != 0 || cs == 0) {
        n.DestroyProgram(&prog);
        log += " (no cubin)";
        return GP_EJIT;
    }
    cubin.resize(cs);
    n.GetCUBIN(prog, cubin.data());
    n.DestroyProgram(&prog);
    return GP_OK;
}

}  // namespace gp
