/*
 * tgb_ops.h -- per-precision launch table. tgb_step_f32.cu and tgb_step_f64.cu each instantiate every kernel for
 * one scalar type and export one table; the C ABI (tgb_capi.cu) picks the table from tgb_world_cfg.precision.
 * The double instantiation is compiled with --fmad=false so that it evaluates the same expression tree as the
 * CPU oracle and can be compared with it at round-off level.
 */
#ifndef TGB_OPS_H
#define TGB_OPS_H

#include <cuda_runtime.h>
#include <stddef.h>
#include "tgb_dev.h"
#include "tgb_io.cuh"
#include "tgb_ccable.cuh"

struct tgb_ops {
    size_t scalar_bytes;
    /* returns cudaErrorInvalidValue for an unsupported lanes-per-env value */
    cudaError_t (*step)(const tgb_step_args& a, int lanes_per_env, int threads, cudaStream_t s);
    cudaError_t (*wide_ctas_per_sm)(const tgb_step_args& a, int lanes_per_env, int* out);
    size_t (*step_smem)(const tgb_step_args& a, int lanes_per_env, int threads);
    cudaError_t (*pack_rods)(const tgb::tgb_io_args& a, const double* src, int broadcast, cudaStream_t s);
    cudaError_t (*unpack_rods)(const tgb::tgb_io_args& a, double* dst, cudaStream_t s);
    cudaError_t (*pack_cables)(const tgb::tgb_io_args& a, const double* src, int broadcast, int reset_motor, cudaStream_t s);
    cudaError_t (*cable_obs)(const tgb::tgb_io_args& a, double* dst, cudaStream_t s);
    cudaError_t (*control_input)(const tgb::tgb_io_args& a, const double* input, int broadcast, double dt, cudaStream_t s);
    cudaError_t (*com)(const tgb::tgb_io_args& a, int mode, double* scores, float* scores_f, cudaStream_t s);
    /* contact cables (tgb_ccable.cuh): one step of every contact cable of the environments in the range; observations / anchors */
    cudaError_t (*ccable_step)(const tgb::tgb_ccable_args& a, cudaStream_t s);
    cudaError_t (*ccable_obs)(const tgb::tgb_ccable_args& a, double* obs, int cable, double* anchors, cudaStream_t s);
    cudaError_t (*ccable_control)(const tgb::tgb_ccable_args& a, const double* input, int broadcast, cudaStream_t s);
};

const tgb_ops* tgb_ops_f32();
const tgb_ops* tgb_ops_f64();

#endif
