/*
 * tgb_inst.cuh -- instantiates every kernel for the scalar type TGB_SCALAR and defines the launch table
 * TGB_OPS_NAME. Included exactly once by tgb_step_f32.cu / tgb_step_f64.cu.
 */
#include "tgb_ops.h"
#include "tgb_step.cuh"

namespace {

typedef TGB_SCALAR scalar_t;

template <int LPE, bool CPG>
cudaError_t launch_step_impl(const tgb_step_args& a, int threads, cudaStream_t s) {
    const size_t smem = tgb::tgb_step_smem_bytes(a, LPE, threads, sizeof(scalar_t));
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tgb::tgb_step_kernel<scalar_t, LPE, CPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const int groupsPerBlock = threads / LPE;
    const int grid = (a.n_env + groupsPerBlock - 1) / groupsPerBlock;
    tgb::tgb_step_kernel<scalar_t, LPE, CPG><<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}
template <int LPE>
cudaError_t launch_step(const tgb_step_args& a, int threads, cudaStream_t s) {
    if (a.ctrl_kind == TGB_CTRL_CPG_IMPEDANCE) return launch_step_impl<LPE, true>(a, threads, s);
    return launch_step_impl<LPE, false>(a, threads, s);
}

cudaError_t step(const tgb_step_args& a, int lpe, int threads, cudaStream_t s) {
    if (threads < 32 || threads > 128 || (threads % 32) != 0) return cudaErrorInvalidValue;
    switch (lpe) {
    case 2: return launch_step<2>(a, threads, s);
    case 4: return launch_step<4>(a, threads, s);
    case 8: return launch_step<8>(a, threads, s);
    case 16: return launch_step<16>(a, threads, s);
    case 32: return launch_step<32>(a, threads, s);
    default: return cudaErrorInvalidValue;
    }
}

size_t step_smem(const tgb_step_args& a, int lpe, int threads) { return tgb::tgb_step_smem_bytes(a, lpe, threads, sizeof(scalar_t)); }

inline int blocksFor(int n, int threads) { return (n + threads - 1) / threads; }

cudaError_t pack_rods(const tgb::tgb_io_args& a, const double* src, int broadcast, cudaStream_t s) {
    const int n = a.env_count * a.R;
    if (n <= 0) return cudaSuccess;
    tgb::pack_rods_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, src, broadcast);
    return cudaGetLastError();
}
cudaError_t unpack_rods(const tgb::tgb_io_args& a, double* dst, cudaStream_t s) {
    const int n = a.env_count * a.R;
    if (n <= 0) return cudaSuccess;
    tgb::unpack_rods_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, dst);
    return cudaGetLastError();
}
cudaError_t pack_cables(const tgb::tgb_io_args& a, const double* src, int broadcast, int reset_motor, cudaStream_t s) {
    const int n = a.env_count * a.C;
    if (n <= 0) return cudaSuccess;
    tgb::pack_cables_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, src, broadcast, reset_motor);
    return cudaGetLastError();
}
cudaError_t cable_obs(const tgb::tgb_io_args& a, double* dst, cudaStream_t s) {
    const int n = a.env_count * a.C;
    if (n <= 0) return cudaSuccess;
    tgb::cable_obs_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, dst);
    return cudaGetLastError();
}
cudaError_t control_input(const tgb::tgb_io_args& a, const double* input, int broadcast, double dt, cudaStream_t s) {
    const int n = a.env_count * a.C;
    if (n <= 0) return cudaSuccess;
    tgb::control_input_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, input, broadcast, dt);
    return cudaGetLastError();
}
cudaError_t com(const tgb::tgb_io_args& a, int mode, double* scores, float* scores_f, cudaStream_t s) {
    if (a.env_count <= 0) return cudaSuccess;
    tgb::com_kernel<scalar_t><<<blocksFor(a.env_count, 128), 128, 0, s>>>(a, mode, scores, scores_f);
    return cudaGetLastError();
}

const tgb_ops kOps = {sizeof(scalar_t), step, step_smem, pack_rods, unpack_rods, pack_cables, cable_obs, control_input, com};

} // namespace

const tgb_ops* TGB_OPS_NAME() { return &kOps; }
