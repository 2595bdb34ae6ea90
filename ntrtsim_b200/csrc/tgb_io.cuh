/*
 * tgb_io.cuh -- kernels on either side of the step: conversion between the host-facing double records of the C ABI
 * (include/ntrtsim_b200.h) and the device state, cable observations, scores and the one-off control input.
 *
 *   rod records     tgb_set_rod_state / tgb_get_rod_state   = tgBaseRigid::centerOfMass / orientation
 *                   (src/core/tgBaseRigid.cpp:74-103) over all bodies of all environments
 *   cable records   tgb_get_cable_obs = tgSpringCable::getRestLength (src/core/tgSpringCable.cpp:65-68),
 *                   tgBulletSpringCable::getActualLength / getTension (src/core/tgBulletSpringCable.cpp:143-155)
 *   control input   tgBasicActuator::setControlInput(input, dt) -> moveMotors (src/core/tgBasicActuator.cpp:149-225)
 *   scores          displacement of the mass centre in the XZ plane
 *                   (src/examples/craterEscape/EscapeController.cpp:119-135,
 *                    src/examples/learningSpines/BaseSpineCPGControl.cpp:281-342)
 */
#pragma once
#include "../../include/ntrtsim_b200.h"
#include "tgb_dev.h"
#include "tgb_math.cuh"

namespace tgb {

struct tgb_io_args {
    void* state;
    double* frame;
    double* com0;
    const void* model_t;
    const int* model_i;
    int n_env, R, C, S, stride, frame_body;
    int frame_static; /* the frame body is a static body: the frame never moves (V = 0) */
    int score_body;   /* >= 0: scores follow this body's centre of mass (CPG controller), -1: the mass centre of all bodies */
    const int* status; /* with score_body >= 0: status bit1 (height window left) turns the score into -1 */
    int env_begin, env_count;
};

/* host doubles [env_count or 1][R][13] -> state + frame. One thread per (env, body). */
template <typename T>
__global__ void pack_rods_kernel(tgb_io_args a, const double* __restrict__ src, int broadcast) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.R) return;
    const int e = i / a.R, b = i - e * a.R;
    const int env = a.env_begin + e;
    const double* rec = src + (size_t)(broadcast ? 0 : e) * a.R * TGB_NBS;
    const double* s = rec + b * TGB_NBS;
    const double* f = rec + a.frame_body * TGB_NBS;
    const double fv[3] = {a.frame_static ? 0.0 : f[7], a.frame_static ? 0.0 : f[8], a.frame_static ? 0.0 : f[9]};
    T* d = reinterpret_cast<T*>(a.state) + (size_t)env * a.stride + b * TGB_NBS;
    d[0] = (T)(s[0] - f[0]); d[1] = (T)(s[1] - f[1]); d[2] = (T)(s[2] - f[2]);
    d[3] = (T)s[3]; d[4] = (T)s[4]; d[5] = (T)s[5]; d[6] = (T)s[6];
    d[7] = (T)(s[7] - fv[0]); d[8] = (T)(s[8] - fv[1]); d[9] = (T)(s[9] - fv[2]);
    d[10] = (T)s[10]; d[11] = (T)s[11]; d[12] = (T)s[12];
    if (b == 0) {
        double* fr = a.frame + (size_t)env * TGB_NFR;
        fr[FR_O] = f[0]; fr[FR_O + 1] = f[1]; fr[FR_O + 2] = f[2];
        fr[FR_V] = fv[0]; fr[FR_V + 1] = fv[1]; fr[FR_V + 2] = fv[2];
    }
}

template <typename T>
__global__ void unpack_rods_kernel(tgb_io_args a, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.R) return;
    const int e = i / a.R, b = i - e * a.R;
    const int env = a.env_begin + e;
    const T* s = reinterpret_cast<const T*>(a.state) + (size_t)env * a.stride + b * TGB_NBS;
    const double* fr = a.frame + (size_t)env * TGB_NFR;
    double* d = dst + ((size_t)e * a.R + b) * TGB_NBS;
    d[0] = fr[FR_O] + (double)s[0]; d[1] = fr[FR_O + 1] + (double)s[1]; d[2] = fr[FR_O + 2] + (double)s[2];
    d[3] = (double)s[3]; d[4] = (double)s[4]; d[5] = (double)s[5]; d[6] = (double)s[6];
    d[7] = fr[FR_V] + (double)s[7]; d[8] = fr[FR_V + 1] + (double)s[8]; d[9] = fr[FR_V + 2] + (double)s[9];
    d[10] = (double)s[10]; d[11] = (double)s[11]; d[12] = (double)s[12];
}

/* host doubles [env_count or 1][C][2] (rest, prev) -> cable state; preferred := rest. One thread per (env, cable). */
template <typename T>
__global__ void pack_cables_kernel(tgb_io_args a, const double* __restrict__ src, int broadcast, int reset_motor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.C) return;
    const int e = i / a.C, c = i - e * a.C;
    const int env = a.env_begin + e;
    const double* s = src + ((size_t)(broadcast ? 0 : e) * a.C + c) * 2;
    T* d = reinterpret_cast<T*>(a.state) + (size_t)env * a.stride + TGB_CABLE_BASE(a.R) + c * TGB_NCS;
    d[CS_REST] = (T)s[0];
    d[CS_PREV] = (T)s[1];
    d[CS_PREF] = (T)s[0];
    if (reset_motor) d[CS_MOTORVEL] = T(0);
}

template <typename T>
struct CableGeom { T L; };

template <typename T>
TGB_D T cable_length(const tgb_io_args& a, const T* rec, int c) {
    const T* mt = reinterpret_cast<const T*>(a.model_t);
    const T* ct = mt + TGB_NBF * a.R + TGB_NSF * a.S;
    const int* ci = a.model_i + TGB_NBI * a.R;
    const int ia = ci[CI_A * a.C + c], ib = ci[CI_B * a.C + c];
    const T* sa = rec + ia * TGB_NBS;
    const T* sb = rec + ib * TGB_NBS;
    const M3<T> Ba = M3<T>::fromQuat(Q4<T>(sa[3], sa[4], sa[5], sa[6]));
    const M3<T> Bb = M3<T>::fromQuat(Q4<T>(sb[3], sb[4], sb[5], sb[6]));
    const V3<T> la(ct[CF_LAX * a.C + c], ct[CF_LAY * a.C + c], ct[CF_LAZ * a.C + c]);
    const V3<T> lb(ct[CF_LBX * a.C + c], ct[CF_LBY * a.C + c], ct[CF_LBZ * a.C + c]);
    const V3<T> p1 = Ba * la + V3<T>(sa[0], sa[1], sa[2]);
    const V3<T> p2 = Bb * lb + V3<T>(sb[0], sb[1], sb[2]);
    return (p2 - p1).length();
}

/* cable observations [env_count][C][4] doubles: rest, length, tension (no damping term), motor velocity */
template <typename T>
__global__ void cable_obs_kernel(tgb_io_args a, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.C) return;
    const int e = i / a.C, c = i - e * a.C;
    const int env = a.env_begin + e;
    const T* rec = reinterpret_cast<const T*>(a.state) + (size_t)env * a.stride;
    const T* cs = rec + TGB_CABLE_BASE(a.R) + c * TGB_NCS;
    const T* mt = reinterpret_cast<const T*>(a.model_t);
    const T k = (mt + TGB_NBF * a.R + TGB_NSF * a.S)[CF_K * a.C + c];
    const T L = cable_length<T>(a, rec, c);
    T tension = (L - cs[CS_REST]) * k;
    tension = tension < T(0) ? T(0) : tension;
    double* d = dst + ((size_t)e * a.C + c) * 4;
    d[0] = (double)cs[CS_REST];
    d[1] = (double)L;
    d[2] = (double)tension;
    d[3] = (double)cs[CS_MOTORVEL];
}

/* tgBasicActuator::setControlInput(input, dt) for every basic actuator; NaN input = leave the cable alone */
template <typename T>
__global__ void control_input_kernel(tgb_io_args a, const double* __restrict__ input, int broadcast, double dt_d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.C) return;
    const int e = i / a.C, c = i - e * a.C;
    const int env = a.env_begin + e;
    const double in = input[(size_t)(broadcast ? 0 : e) * a.C + c];
    if (!(in >= 0.0)) return;
    const int* ci = a.model_i + TGB_NBI * a.R;
    if (ci[CI_ACTUATOR * a.C + c] != TGB_ACT_BASIC) return;
    T* rec = reinterpret_cast<T*>(a.state) + (size_t)env * a.stride;
    T* cs = rec + TGB_CABLE_BASE(a.R) + c * TGB_NCS;
    const T* ct = reinterpret_cast<const T*>(a.model_t) + TGB_NBF * a.R + TGB_NSF * a.S;
    const T dt = (T)dt_d;
    const T k = ct[CF_K * a.C + c];
    const T L = cable_length<T>(a, rec, c);
    T pref = (T)in, rest = cs[CS_REST];
    const T stepSize = ct[CF_TVEL * a.C + c] * dt;
    const T maxTens = ct[CF_MAXTENS * a.C + c];
    if ((L - pref) * k > maxTens) pref = L - ct[CF_MAXSTRETCH * a.C + c];
    const T diff = pref - rest;
    const T fabsDiff = tabs(diff);
    if ((L > ct[CF_MINAL * a.C + c]) || (diff > T(0))) {
        if (fabsDiff > stepSize) rest += (diff / fabsDiff) * stepSize;
        else rest += diff;
    }
    const T minRL = ct[CF_MINRL * a.C + c];
    rest = (rest > minRL) ? rest : minRL;
    cs[CS_REST] = rest;
    cs[CS_PREF] = pref;
}

/* mass-weighted centre of mass per env, in double. mode 0: store into com0; mode 1: scores[env][2] (double) and
   scores_f[env][2] (float) = { XZ distance from com0, 0 } */
template <typename T>
__global__ void com_kernel(tgb_io_args a, int mode, double* __restrict__ scores, float* __restrict__ scores_f) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.env_count) return;
    const int env = a.env_begin + e;
    const T* rec = reinterpret_cast<const T*>(a.state) + (size_t)env * a.stride;
    const T* mt = reinterpret_cast<const T*>(a.model_t);
    const double* fr = a.frame + (size_t)env * TGB_NFR;
    double sx = 0, sy = 0, sz = 0, m = 0;
    for (int b = 0; b < a.R; ++b) {
        const double mass = (double)mt[BF_MASS * a.R + b];
        const T* s = rec + b * TGB_NBS;
        sx += (fr[FR_O] + (double)s[0]) * mass;
        sy += (fr[FR_O + 1] + (double)s[1]) * mass;
        sz += (fr[FR_O + 2] + (double)s[2]) * mass;
        m += mass;
    }
    if (m > 0) { const double im = 1.0 / m; sx *= im; sy *= im; sz *= im; }
    if (a.score_body >= 0) {
        const T* s = rec + a.score_body * TGB_NBS;
        sx = fr[FR_O] + (double)s[0]; sy = fr[FR_O + 1] + (double)s[1]; sz = fr[FR_O + 2] + (double)s[2];
    }
    double* c0 = a.com0 + (size_t)env * 3;
    if (mode == 0) {
        c0[0] = sx; c0[1] = sy; c0[2] = sz;
    } else {
        const double dx = sx - c0[0], dz = sz - c0[2];
        double d = sqrt(dx * dx + dz * dz);
        if (a.score_body >= 0 && a.status != nullptr && (a.status[env] & TGB_STATUS_HEIGHT)) d = -1.0; /* bogus episode */
        if (scores) { scores[(size_t)e * 2] = d; scores[(size_t)e * 2 + 1] = 0.0; }
        if (scores_f) { scores_f[(size_t)env * 2] = (float)d; scores_f[(size_t)env * 2 + 1] = 0.0f; }
    }
}

} // namespace tgb
