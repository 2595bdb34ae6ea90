/*
 * tgb_math.cuh -- small fixed-size vector / matrix / quaternion helpers for the batched step kernels.
 * Arithmetic order follows Bullet's LinearMath as the reference uses it (btVector3 operator/ multiplies by the
 * reciprocal, btMatrix3x3::setRotation / scaled / transpose, btQuaternion product), so that the double-precision
 * instantiation can be compared with the CPU oracle at round-off level.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define TGB_D __device__ __forceinline__
#define TGB_HOST_DEVICE_INLINE __host__ __device__ inline

namespace tgb {

TGB_D float tsqrt(float x) { return sqrtf(x); }
TGB_D double tsqrt(double x) { return sqrt(x); }
/* Square root without the library routine's special-case branch (it ends a basic block, which keeps ptxas from overlapping the
   dependent chains of independent cables): MUFU.RSQ, one Newton step, zero handled by a select; within 1 ulp of sqrtf. */
TGB_D float tsqrtNoBranch(float x) {
    const float r = rsqrtf(x);
    const float s = x * r;
    const float t = fmaf(fmaf(-s, s, x), 0.5f * r, s);
    return x > 0.0f ? t : 0.0f;
}
TGB_D double tsqrtNoBranch(double x) { return sqrt(x); }
TGB_D float tabs(float x) { return fabsf(x); }
TGB_D double tabs(double x) { return fabs(x); }
TGB_D float tsin(float x) { return sinf(x); }
TGB_D double tsin(double x) { return sin(x); }
TGB_D float tcos(float x) { return cosf(x); }
TGB_D double tcos(double x) { return cos(x); }
/* reciprocal / reciprocal square root / copysign: exact IEEE forms in double (the parity instrument), the fast hardware
   approximations (MUFU + refinement, <= 2 ulp) in float (the throughput configuration) */
TGB_D float trcp(float x) { return __fdividef(1.0f, x); }
TGB_D double trcp(double x) { return 1.0 / x; }
TGB_D float tdiv(float a, float b) { return __fdividef(a, b); }
TGB_D double tdiv(double a, double b) { return a / b; }
TGB_D float trsqrt(float x) { return rsqrtf(x); }
TGB_D double trsqrt(double x) { return 1.0 / sqrt(x); }
TGB_D float tcopysign(float a, float b) { return copysignf(a, b); }
TGB_D double tcopysign(double a, double b) { return copysign(a, b); }

template <typename T>
struct V3 {
    T x, y, z;
    TGB_D V3() {}
    TGB_D V3(T a, T b, T c) : x(a), y(b), z(c) {}
    TGB_D V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    TGB_D V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    TGB_D V3 operator-() const { return V3(-x, -y, -z); }
    TGB_D V3 operator*(T s) const { return V3(x * s, y * s, z * s); }
    TGB_D V3 operator/(T s) const { T i = T(1) / s; return V3(x * i, y * i, z * i); }
    TGB_D V3& operator+=(const V3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    TGB_D V3& operator-=(const V3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    TGB_D V3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
    TGB_D T dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    TGB_D V3 cross(const V3& o) const { return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    TGB_D T length2() const { return dot(*this); }
    TGB_D T length() const { return tsqrt(length2()); }
    TGB_D bool nonzero() const { return x != T(0) || y != T(0) || z != T(0); }
};

template <typename T>
struct Q4 {
    T x, y, z, w;
    TGB_D Q4() {}
    TGB_D Q4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
};

/* btQuaternion operator*(q1, q2) */
template <typename T>
TGB_D Q4<T> qmul(const Q4<T>& a, const Q4<T>& b) {
    return Q4<T>(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                 a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
                 a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
                 a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}

template <typename T>
struct M3 {
    T m[3][3];
    TGB_D V3<T> row(int i) const { return V3<T>(m[i][0], m[i][1], m[i][2]); }
    TGB_D V3<T> col(int j) const { return V3<T>(m[0][j], m[1][j], m[2][j]); }
    TGB_D V3<T> operator*(const V3<T>& v) const { return V3<T>(row(0).dot(v), row(1).dot(v), row(2).dot(v)); }
    /* v * M (row vector times matrix) */
    TGB_D V3<T> tmul(const V3<T>& v) const { return V3<T>(col(0).dot(v), col(1).dot(v), col(2).dot(v)); }
    /* btMatrix3x3::setRotation */
    TGB_D static M3 fromQuat(const Q4<T>& q) {
        T d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        T s = tdiv(T(2), d);
        T xs = q.x * s, ys = q.y * s, zs = q.z * s;
        T wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
        T xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
        T yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
        M3 r;
        r.m[0][0] = T(1) - (yy + zz); r.m[0][1] = xy - wz;          r.m[0][2] = xz + wy;
        r.m[1][0] = xy + wz;          r.m[1][1] = T(1) - (xx + zz); r.m[1][2] = yz - wx;
        r.m[2][0] = xz - wy;          r.m[2][1] = yz + wx;          r.m[2][2] = T(1) - (xx + yy);
        return r;
    }
    /* (this.scaled(s)) * this^T = R diag(s) R^T, evaluated like btMatrix3x3 scaled() then operator* */
    TGB_D M3 conjugateDiag(const V3<T>& s) const {
        M3 sc;
#pragma unroll
        for (int i = 0; i < 3; ++i) { sc.m[i][0] = m[i][0] * s.x; sc.m[i][1] = m[i][1] * s.y; sc.m[i][2] = m[i][2] * s.z; }
        M3 r;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) r.m[i][j] = sc.m[i][0] * m[j][0] + sc.m[i][1] * m[j][1] + sc.m[i][2] * m[j][2];
        return r;
    }
};

} // namespace tgb
