// Kernels either side of the solve for scenario studies:
//
//   gp_der_rows      synthesis of the packed spu rows of an hourly load profile x DER scenario
//                    study on the device (BASELINE config C4): the inputs of a million-scenario
//                    study are two hour tables and a DER table of a few kilobytes, not gigabytes
//                    of per-scenario loads crossing PCIe;
//   gp_vmag_summary  per-scenario min |V|, max |V| and count of bus-phases outside a voltage band,
//                    so that a study returns 20 bytes per scenario instead of P_b doubles.
//
// Both are streaming, HBM-bound kernels: one thread per scenario, rows walked with the scenario
// index innermost (coalesced), constants broadcast through L1.
#include "gp_common.cuh"

namespace gp {

// spu[row][s] = m_h[h] * W0[row] - (Re W0[row] * mask[d][row]) * (pen[d] * pv_h[h]),  scenario
// index first + s = d * H + h.  The products and the difference are rounded separately
// (__dmul_rn / __dsub_rn, no FMA contraction) so that the rows are bit-identical to the NumPy
// expression the workload is defined by (bench.py:make_der_rows).
__global__ void __launch_bounds__(256)
der_rows_kernel(long long S, long long ld, long long first, long long stride, int H, int D, int n_srows,
                const double2* __restrict__ W0, const double* __restrict__ m_h, const double* __restrict__ pv_h,
                const double* __restrict__ mask, const double* __restrict__ pen, double2* __restrict__ spu) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const long long idx = (first + s * stride) % ((long long)H * D);
    const int d = (int)(idx / H), h = (int)(idx % H);
    const double m = __ldg(m_h + h);
    const double g = __dmul_rn(__ldg(pen + d), __ldg(pv_h + h));
    for (int r = blockIdx.y; r < n_srows; r += gridDim.y) {
        const double2 w = __ldg(W0 + r);
        const double pvr = __dmul_rn(__dmul_rn(w.x, __ldg(mask + (long long)d * n_srows + r)), g);
        spu[(long long)r * ld + s] = make_double2(__dsub_rn(__dmul_rn(w.x, m), pvr), __dmul_rn(w.y, m));
    }
}

__global__ void __launch_bounds__(256)
vmag_summary_kernel(long long S, long long ld, int n_vrows, const double* __restrict__ Vmag, double vmin,
                    double vmax, double* __restrict__ out_min, double* __restrict__ out_max,
                    int* __restrict__ out_count) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = -lo;
    int cnt = 0;
    for (int r = 0; r < n_vrows; ++r) {
        const double a = Vmag[(long long)r * ld + s];
        lo = fmin(lo, a);
        hi = fmax(hi, a);
        cnt += !(a >= vmin && a <= vmax);       // NaN counts as a violation
    }
    if (out_min) out_min[s] = lo;
    if (out_max) out_max[s] = hi;
    if (out_count) out_count[s] = cnt;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_der_rows(int device, int64_t S, int64_t ld, int64_t first, int32_t H, int32_t D, int32_t n_srows,
                           const double* W0_dev, const double* m_h_dev, const double* pv_h_dev,
                           const double* mask_dev, const double* pen_dev, double* spu_dev, void* stream) {
    return gp_der_rows_strided(device, S, ld, first, 1, H, D, n_srows, W0_dev, m_h_dev, pv_h_dev, mask_dev, pen_dev,
                               spu_dev, stream);
}

extern "C" int gp_der_rows_strided(int device, int64_t S, int64_t ld, int64_t first, int64_t stride, int32_t H,
                                   int32_t D, int32_t n_srows, const double* W0_dev, const double* m_h_dev,
                                   const double* pv_h_dev, const double* mask_dev, const double* pen_dev,
                                   double* spu_dev, void* stream) {
    if (S < 0 || ld < S || H <= 0 || D <= 0 || n_srows < 0 || first < 0 || stride < 1)
        return fail(GP_EINVAL, "gp_der_rows: bad sizes");
    if (S == 0 || n_srows == 0) return GP_OK;
    if (!W0_dev || !m_h_dev || !pv_h_dev || !mask_dev || !pen_dev || !spu_dev)
        return fail(GP_EINVAL, "gp_der_rows: null buffer");
    GP_CUDA(cudaSetDevice(device));
    dim3 grid((unsigned)((S + 255) / 256), (unsigned)(n_srows < 8 ? n_srows : 8));
    der_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S, ld, first, stride, H, D, n_srows, (const double2*)W0_dev,
                                                            m_h_dev, pv_h_dev, mask_dev, pen_dev, (double2*)spu_dev);
    count_launch();
    GP_CUDA(cudaGetLastError());
    return GP_OK;
}

extern "C" int gp_vmag_summary(int device, int64_t S, int64_t ld, int32_t n_vrows, const double* Vmag_dev,
                               double vmin, double vmax, double* min_dev, double* max_dev, int32_t* count_dev,
                               void* stream) {
    if (S < 0 || ld < S || n_vrows < 0) return fail(GP_EINVAL, "gp_vmag_summary: bad sizes");
    if (S == 0) return GP_OK;
    if (!Vmag_dev) return fail(GP_EINVAL, "gp_vmag_summary: null buffer");
    GP_CUDA(cudaSetDevice(device));
    vmag_summary_kernel<<<(unsigned)((S + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        S, ld, n_vrows, Vmag_dev, vmin, vmax, min_dev, max_dev, count_dev);
    count_launch();
    GP_CUDA(cudaGetLastError());
    return GP_OK;
}
