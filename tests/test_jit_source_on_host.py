"""The feeder-specialised FBS kernel is GENERATED source (gp_fbs_jit.cu).  Its sweep function is
nearly plain C++, so this test compiles the generated text for the HOST (g++, a few shims for
double2 / __ldcg / the rsqrt seed), runs one lane of it scenario by scenario exactly as the kernel's
solve loop does, and compares V, I and the iteration counts with the oracle.  It checks the
generator's schedule, row offsets, constants and the shared I slots (fbs_jit_islots) without a GPU;
the GPU tests then check the same source as compiled by NVRTC.  Test infrastructure only: nothing
in the product runs on the host."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from helpers import ALL_FEEDERS, ZIPS, load_model

SHIMS = r'''
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __constant__ static const
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
using std::isinf;
'''

MAIN = r'''
int main(int argc, char** argv) {
    // argv: S tol maxiter in.bin out.bin ; in = spu[NS][S] complex ; out = V[NV][S], I[NI][S] complex, it[S] double
    const long S = atol(argv[1]); const double tol = atof(argv[2]); const int maxiter = atoi(argv[3]);
    std::vector<double2> spu((size_t)(NS > 0 ? NS : 1) * S), V((size_t)NV * S), I((size_t)(NI > 0 ? NI : 1) * S);
    std::vector<double> its(S);
    FILE* f = fopen(argv[4], "rb"); if (!f) return 2;
    if (NS > 0 && fread(spu.data(), 16, (size_t)NS * S, f) != (size_t)NS * S) return 3;
    fclose(f);
    const unsigned ldb = (unsigned)(S * 16);
    const double2 zero = make_double2(0.0, 0.0);
    for (long s = 0; s < S; ++s) {
        std::vector<double2> sm((size_t)SROWS * 32 + 32, zero);
        double2* const Vs = sm.data();
#if IG_MODE
        IPTR const Ip = reinterpret_cast<char*>(I.data() + s);      // I rows in the caller's array, [row][S]
#else
        IPTR const Ip = Vs + NV * 32;
#endif
        const char* sp = reinterpret_cast<const char*>(spu.data() + s);
        for (int r = 0; r < NV; ++r) Vs[r * 32] = zero;
        for (int r = 0; r < NIP; ++r) IST(r, zero);
        int it = 0; double diff = -1.0; bool converged = false, exact = false;
        while (!converged && it < maxiter) {
            bool bad = false;
            if (exact) diff = gp_sweep<true>(Vs, Ip, sp, ldb, bad);
            else diff = gp_sweep<false>(Vs, Ip, sp, ldb, bad);
            if (bad) {
                for (int r = 0; r < NV; ++r) Vs[r * 32] = zero;
                for (int r = 0; r < NIP; ++r) IST(r, zero);
                it = 0; exact = true; continue;
            }
            ++it; converged = diff <= tol;
        }
        for (int r = 0; r < NV; ++r) V[(size_t)r * S + s] = Vs[r * 32];
#if !IG_MODE
        char* Ig = reinterpret_cast<char*>(I.data() + s);
        GP_I_ROWS_OUT(Ig)
#endif
        its[s] = it;
    }
    f = fopen(argv[5], "wb"); if (!f) return 4;
    fwrite(V.data(), 16, (size_t)NV * S, f);
    fwrite(I.data(), 16, (size_t)(NI > 0 ? NI : 1) * S, f);
    fwrite(its.data(), 8, S, f);
    fclose(f);
    return 0;
}
'''


def host_program(src: str) -> str:
    head = src[:src.index('extern "C" __global__')]
    lines = []
    for ln in head.splitlines():
        if 'rsqrt.approx' in ln:      # the PTX seed (2^-22 accurate): any seed of that quality converges the same.
            # (mantissa cut to 23 bits in place, not a cast to float: the cast loses the RANGE too, and a diverging sweep
            # - |V| of 1e45 and more with heavy constant-impedance loads, fuzzer seed 81081 - got a zero seed from it)
            ln = ('    { double s_ = 1.0 / std::sqrt(m2); unsigned long long b_; memcpy(&b_, &s_, 8); '
                  'b_ &= 0xFFFFFFFFE0000000ULL; memcpy(&y, &b_, 8); }')
        lines.append(ln)
    return SHIMS + '\n'.join(lines) + '\n' + MAIN


@pytest.mark.parametrize('ig', [0, 1])
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_generated_sweep_run_on_the_host_matches_the_oracle(feeder, ig, tmp_path):
    from gigpower_b200 import _cabi
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs
    from oracle.gp_oracle import FBSOracle
    model = load_model(feeder)
    t = flatten_fbs(Circuit(model))
    d, keep = _cabi.make_fbs_desc(t)
    n = _cabi.gp_fbs_jit_source(C.byref(d), 2, ig, None, 0)      # ig = 1: I rows in global memory
    buf = C.create_string_buffer(n + 1)
    assert _cabi.gp_fbs_jit_source(C.byref(d), 2, ig, buf, n + 1) == n
    src = buf.value.decode()
    nip = int([ln for ln in src.splitlines() if ln.startswith('#define NIP ')][0].split()[-1])
    assert 0 < nip <= t.n_irows and (nip == t.n_irows if ig else True)
    cpp, exe = tmp_path / 'sweep.cpp', tmp_path / 'sweep'
    cpp.write_text(host_program(src))
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', str(exe), str(cpp)], check=True)

    o = FBSOracle(model, ZIPS['test_zip'])
    rng = np.random.Generator(np.random.PCG64(99))
    S = 12
    mult = rng.uniform(0.4, 1.6, size=(S, len(o.c.load_names)))
    mult[0] = 1.0
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)          # [S, nb, 3]
    want = o.solve(spu)
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), S), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    fin, fout = tmp_path / 'in.bin', tmp_path / 'out.bin'
    rows[:t.n_srows].astype(np.complex128).tofile(fin)
    subprocess.run([str(exe), str(S), repr(float(t.tolerance)), '100', str(fin), str(fout)], check=True)
    raw = np.fromfile(fout, dtype=np.float64)
    nv, ni = t.n_vrows, max(t.n_irows, 1)
    Vr = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
    Ir = raw[2 * nv * S:2 * (nv + ni) * S].view(np.complex128).reshape(ni, S)
    its = raw[2 * (nv + ni) * S:].astype(int)
    assert list(its) == list(want['iterations'])
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(Vr[t.vrow[bb, pp]].T - want['V'][:, bb, pp]).max() <= 1e-9
    ll, qq = np.nonzero(t.irow >= 0)
    assert np.abs(Ir[t.irow[ll, qq]].T - want['I'][:, ll, qq]).max() <= 1e-9


@pytest.mark.parametrize('feeder,n_irows,n_slots', [('IEEE_13_Bus_allwye_noxfm_noreg', 35, 22), ('IEEE_13_Bus_allwye', 35, 22),
                                                    ('IEEE_34_Bus_allwye_noxfm_noreg', 86, 55), ('IEEE_34_Bus_allwye', 86, 55),
                                                    ('IEEE_37_Bus_allwye_noxfm_noreg', 114, 77), ('IEEE_37_Bus_allwye', 111, 77)])
def test_shared_i_slots_per_feeder(feeder, n_irows, n_slots):
    """fbs_jit_islots: one shared-memory slot per distinct I value (rows that repeat a child's current share
    its slot, rows that stay zero take none); with I rows in global memory every row keeps its own."""
    from gigpower_b200 import _cabi
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs
    t = flatten_fbs(Circuit(load_model(feeder)))
    d, keep = _cabi.make_fbs_desc(t)
    for ig, want in ((0, n_slots), (1, n_irows)):
        n = _cabi.gp_fbs_jit_source(C.byref(d), 2, ig, None, 0)
        buf = C.create_string_buffer(n + 1)
        assert _cabi.gp_fbs_jit_source(C.byref(d), 2, ig, buf, n + 1) == n
        defs = dict(ln.split()[1:3] for ln in buf.value.decode().splitlines() if ln.startswith('#define N'))
        assert int(defs['NI']) == t.n_irows == n_irows and int(defs['NIP']) == want


MAIN_EX = r'''
int main(int argc, char** argv) {
    // argv: S tol maxiter in.bin gam.bin out.bin ; in = spu[NS][S] complex ; gam = gamma[NG][S] double
    const long S = atol(argv[1]); const double tol = atof(argv[2]); const int maxiter = atoi(argv[3]); const int NG = atoi(argv[7]);
    std::vector<double2> spu((size_t)(NS > 0 ? NS : 1) * S), V((size_t)NV * S), I((size_t)(NI > 0 ? NI : 1) * S);
    std::vector<double> gam((size_t)(NG > 0 ? NG : 1) * S), its(S);
    FILE* f = fopen(argv[4], "rb"); if (!f) return 2;
    if (NS > 0 && fread(spu.data(), 16, (size_t)NS * S, f) != (size_t)NS * S) return 3;
    fclose(f);
    f = fopen(argv[5], "rb"); if (!f) return 2;
    if (NG > 0 && fread(gam.data(), 8, (size_t)NG * S, f) != (size_t)NG * S) return 3;
    fclose(f);
    const unsigned ldb = (unsigned)(S * 16);
    const double2 zero = make_double2(0.0, 0.0);
    for (long s = 0; s < S; ++s) {
        std::vector<double2> sm((size_t)SROWS * 32 + 32, zero);
        double2* const Vs = sm.data();
        IPTR const Ip = Vs + NV * 32;
        const char* sp = reinterpret_cast<const char*>(spu.data() + s);
        const char* gp = reinterpret_cast<const char*>(gam.data() + s);
        int it = 0; double diff = -1.0; bool converged = false, exact = false;
        while (!converged && it < maxiter) {
            bool bad = false;
            if (exact) diff = gp_sweep<true>(Vs, Ip, sp, ldb, bad, gp);
            else diff = gp_sweep<false>(Vs, Ip, sp, ldb, bad, gp);
            if (bad) {
                for (int r = 0; r < NV; ++r) Vs[r * 32] = zero;
                for (int r = 0; r < NIP; ++r) IST(r, zero);
                it = 0; exact = true; continue;
            }
            ++it; converged = diff <= tol;
        }
        for (int r = 0; r < NV; ++r) V[(size_t)r * S + s] = Vs[r * 32];
        char* Ig = reinterpret_cast<char*>(I.data() + s);
        GP_I_ROWS_OUT(Ig)
        its[s] = it;
    }
    f = fopen(argv[6], "wb"); if (!f) return 4;
    fwrite(V.data(), 16, (size_t)NV * S, f);
    fwrite(I.data(), 16, (size_t)(NI > 0 ? NI : 1) * S, f);
    fwrite(its.data(), 8, S, f);
    fclose(f);
    return 0;
}
'''


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye', 'IEEE_34_Bus_allwye', 'IEEE_37_Bus_allwye'])
def test_generated_sweep_with_per_scenario_taps_matches_the_oracle(feeder, tmp_path):
    """The EX build of the specialised kernel (gp_fbs_jit_source mode + 2) reads every regulated phase's ratio from
    the scenario's column of the gamma rows: compiled for the host and run scenario by scenario, it must give what
    one oracle per tap vector gives (oracle.gp_oracle.solve_with_taps, itself pinned to one reference object per
    scenario in tests/test_oracle_golden.py)."""
    from gigpower_b200 import _cabi
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs
    from oracle.gp_oracle import FBSOracle, solve_with_taps
    model = load_model(feeder)
    c = Circuit(model)
    t = flatten_fbs(c)
    d, keep = _cabi.make_fbs_desc(t)
    n = _cabi.gp_fbs_jit_source(C.byref(d), 2, 2, None, 0)
    buf = C.create_string_buffer(n + 1)
    assert _cabi.gp_fbs_jit_source(C.byref(d), 2, 2, buf, n + 1) == n
    src = buf.value.decode()
    assert '#define EX_MODE 1' in src
    prog = host_program(src)
    prog = prog[:prog.index('int main(')] + MAIN_EX
    cpp, exe = tmp_path / 'sweep_ex.cpp', tmp_path / 'sweep_ex'
    cpp.write_text(prog)
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', str(exe), str(cpp)], check=True)
    o = FBSOracle(model, ZIPS['test_zip'])
    rng = np.random.Generator(np.random.PCG64(77))
    S = 10
    mult = rng.uniform(0.5, 1.4, size=(S, len(o.c.load_names)))
    regs = [vr.__name__ for vr in c.voltage_regulators.get_elements()]
    taps = rng.integers(-16, 17, size=(S, len(regs)))
    taps[0] = 16
    taps[1] = -16
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    assert regs == [rc.name for rc in model.regcontrols]      # solve_with_taps takes the taps in this order
    want = solve_with_taps(FBSOracle, model, spu, taps, ZIPS['test_zip'])
    g = (taps.astype(float) + 16) / 32
    g = g * (1.10 - .90)
    g = g + .90                                             # voltage_regulator.py:23-39, same operation order
    gam = np.ascontiguousarray(g.T[np.asarray(t.grow_vr, dtype=int)])       # [n_gamma_rows, S]
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), S), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    fin, fg, fout = tmp_path / 'in.bin', tmp_path / 'gam.bin', tmp_path / 'out.bin'
    rows[:t.n_srows].astype(np.complex128).tofile(fin)
    gam.astype(np.float64).tofile(fg)
    subprocess.run([str(exe), str(S), repr(float(t.tolerance)), '100', str(fin), str(fg), str(fout), str(gam.shape[0])], check=True)
    raw = np.fromfile(fout, dtype=np.float64)
    nv, ni = t.n_vrows, max(t.n_irows, 1)
    Vr = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
    its = raw[2 * (nv + ni) * S:].astype(int)
    assert list(its) == list(want['iterations'])
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(Vr[t.vrow[bb, pp]].T - want['V'][:, bb, pp]).max() <= 1e-9


MAIN_DFS = r'''
int main(int argc, char** argv) {
    // argv: S tol maxiter in.bin out.bin ; the depth-first build: V rows live in the caller's array
    const long S = atol(argv[1]); const double tol = atof(argv[2]); const int maxiter = atoi(argv[3]);
    std::vector<double2> spu((size_t)(NS > 0 ? NS : 1) * S), V((size_t)NV * S), I((size_t)(NI > 0 ? NI : 1) * S);
    std::vector<double> its(S);
    FILE* f = fopen(argv[4], "rb"); if (!f) return 2;
    if (NS > 0 && fread(spu.data(), 16, (size_t)NS * S, f) != (size_t)NS * S) return 3;
    fclose(f);
    const unsigned ldb = (unsigned)(S * 16);
    const double2 zero = make_double2(0.0, 0.0);
    for (long s = 0; s < S; ++s) {
        std::vector<double2> sm((size_t)SROWS * 32 + 32, zero);
        double2* const Vs = sm.data();
        IPTR const Ip = Vs + NVS * 32;
        char* const Vw = reinterpret_cast<char*>(V.data() + s);
        const char* sp = reinterpret_cast<const char*>(spu.data() + s);
        for (int r = 0; r < NIP; ++r) IST(r, zero);
        int it = 0; double diff = -1.0; bool converged = false, exact = false;
        while (!converged && it < maxiter) {
            bool bad = false;
            if (exact) diff = gp_sweep<true>(Vs, Ip, sp, ldb, bad, Vw);
            else diff = gp_sweep<false>(Vs, Ip, sp, ldb, bad, Vw);
            if (bad) {
                for (int r = 0; r < NIP; ++r) IST(r, zero);
                it = 0; exact = true; continue;
            }
            ++it; converged = diff <= tol;
        }
        char* Ig = reinterpret_cast<char*>(I.data() + s);
        GP_I_ROWS_OUT(Ig)
        its[s] = it;
    }
    f = fopen(argv[5], "wb"); if (!f) return 4;
    fwrite(V.data(), 16, (size_t)NV * S, f);
    fwrite(I.data(), 16, (size_t)(NI > 0 ? NI : 1) * S, f);
    fwrite(its.data(), 8, S, f);
    fclose(f);
    return 0;
}
'''


@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_generated_depth_first_sweep_run_on_the_host_matches_the_oracle(feeder, tmp_path):
    """The depth-first build of the specialised kernel (gp_fbs_jit_source mode + 4: V rows in the caller's array, I slots
    and a small stack in shared memory, a bus's forward step on the way down and its backward step on the way up):
    compiled for the host and run scenario by scenario against the oracle - V, I, iteration counts - incl. scenarios
    that take the exact restart path."""
    from gigpower_b200 import _cabi
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs
    from oracle.gp_oracle import FBSOracle
    model = load_model(feeder)
    t = flatten_fbs(Circuit(model))
    d, keep = _cabi.make_fbs_desc(t)
    n = _cabi.gp_fbs_jit_source(C.byref(d), 2, 4, None, 0)
    assert n > 0
    buf = C.create_string_buffer(n + 1)
    assert _cabi.gp_fbs_jit_source(C.byref(d), 2, 4, buf, n + 1) == n
    src = buf.value.decode()
    defs = dict(ln.split()[1:3] for ln in src.splitlines() if ln.startswith('#define N') or ln.startswith('#define DFS'))
    assert defs['DFS_MODE'] == '1' and defs['NVS'] == '0' and 0 < int(defs['NST']) <= 24
    prog = host_program(src)
    prog = prog[:prog.index('int main(')] + MAIN_DFS
    cpp, exe = tmp_path / 'sweep_dfs.cpp', tmp_path / 'sweep_dfs'
    cpp.write_text(prog)
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', str(exe), str(cpp)], check=True)
    o = FBSOracle(model, ZIPS['test_zip'])
    rng = np.random.Generator(np.random.PCG64(98))
    S = 12
    mult = rng.uniform(0.4, 1.6, size=(S, len(o.c.load_names)))
    mult[0] = 1.0
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    want = o.solve(spu)
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), S), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    fin, fout = tmp_path / 'in.bin', tmp_path / 'out.bin'
    rows[:t.n_srows].astype(np.complex128).tofile(fin)
    subprocess.run([str(exe), str(S), repr(float(t.tolerance)), '100', str(fin), str(fout)], check=True)
    raw = np.fromfile(fout, dtype=np.float64)
    nv, ni = t.n_vrows, max(t.n_irows, 1)
    Vr = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
    Ir = raw[2 * nv * S:2 * (nv + ni) * S].view(np.complex128).reshape(ni, S)
    its = raw[2 * (nv + ni) * S:].astype(int)
    assert list(its) == list(want['iterations'])
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(Vr[t.vrow[bb, pp]].T - want['V'][:, bb, pp]).max() <= 1e-9
    ll, qq = np.nonzero(t.irow >= 0)
    assert np.abs(Ir[t.irow[ll, qq]].T - want['I'][:, ll, qq]).max() <= 1e-9
