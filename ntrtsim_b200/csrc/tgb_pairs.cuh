/*
 * tgb_pairs.cuh -- contact between the bodies of one environment (rod x rod, vertebra x vertebra).
 *
 * Reference behaviour replaced: tgWorldBulletPhysicsImpl.cpp:82-107 builds the world with btAxisSweep3 and
 * btDefaultCollisionConfiguration, i.e. every pair of rigid bodies collides (tgRodInfo.cpp:99-113 btCylinderShape,
 * tgSphereInfo.cpp:100-115 btSphereShape), and the constraint solver couples the two bodies of a manifold point.
 *
 * Manifold (same scheme as the CPU oracle's generateBodyPairContacts): a cylinder counts as the capsule inscribed in it (its axis
 * segment shortened by the radius at both ends, swept by the radius), a sphere as a segment of length zero; one contact at the
 * closest points of the two core segments, two (the ends of the overlap) when the segments are parallel within kParallel and
 * overlap; a point is a contact when the surface distance is within the contact breaking threshold; at most 4 rows per pair of
 * bodies and pair_cap rows per environment, in the order (body A < body B, shape of A, shape of B, point).  Boxes take no part
 * (static boxes are obstacles, see tgb_step.cuh).
 *
 * Mapping.  Detection runs on every step: the lanes of an environment's group take the body pairs round robin and test them
 * from the published poses (a cull by bounding spheres, then the closest points of the core segments).  Nothing else happens
 * in the common case of no pair within the contact band.  Otherwise the lanes that found one emit the rows one after the
 * other (row order = pair order, as in the oracle), and the environment's bodies leave their independent per-lane solvers for
 * this step: rows against the static world stay with their body's lane, a row between two bodies is worked on by the two lanes
 * that own the bodies -- each forms its side's Jacobian block from its own registers, the two relative velocities meet through
 * one shuffle -- and every iteration of the projected Gauss-Seidel visits the rows in the oracle's order (static-world rows body by
 * body, then the pair rows; normal rows, then friction rows).
 */
#pragma once
#include "tgb_dev.h"
#include "tgb_math.cuh"

namespace tgb {

template <typename T> struct ContactSolveOut { V3<T> dLin, dAng, push, turn; };

template <typename T> TGB_D T pairBreakingThreshold() { return T(0.02); }
template <typename T> TGB_D T pairParallel() { return T(1e-4); }

template <typename T>
struct CoreSegment {
    V3<T> p, d; /* centre, unit direction */
    T h, r;     /* half length of the core segment, radius of the sweep */
};

template <typename T> TGB_D T clampT(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* core segment of collision shape si of the body whose published pose record (pos3, basis9) is `pose`; false for a box */
template <typename T, typename Model>
TGB_D bool coreSegment(const Model& M, const T* pose, int si, CoreSegment<T>& g) {
    if (M.shape(SF_KIND, si) == T(TGB_SHAPE_BOX)) return false;
    const V3<T> x(pose[0], pose[1], pose[2]);
    M3<T> B;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B.m[i][j] = pose[3 + i * 3 + j];
    const V3<T> spos(M.shape(SF_POSX, si), M.shape(SF_POSY, si), M.shape(SF_POSZ, si));
    const V3<T> r1(M.shape(SF_ROT0 + 1, si), M.shape(SF_ROT0 + 4, si), M.shape(SF_ROT0 + 7, si));
    g.p = B * spos + x;
    g.d = B * r1;
    g.r = M.shape(SF_RADIUS, si);
    g.h = M.shape(SF_SEGH, si);
    return true;
}

/* closest points A(s) = pa + s da, B(t) = pb + t db of two segments given by centre, unit direction and half length */
template <typename T>
TGB_D void closestSegmentParams(const CoreSegment<T>& A, const CoreSegment<T>& B, T& sOut, T& tOut, T& bOut, T& cOut, T& fOut) {
    const V3<T> r = A.p - B.p;
    const T b = A.d.dot(B.d), c = A.d.dot(r), f = B.d.dot(r);
    const T denom = T(1) - b * b;
    T s = T(0), t = T(0);
    if (A.h > T(0) && denom > T(1e-12)) s = clampT((b * f - c) / denom, -A.h, A.h);
    t = b * s + f;
    if (t < -B.h) { t = -B.h; s = clampT(b * t - c, -A.h, A.h); }
    else if (t > B.h) { t = B.h; s = clampT(b * t - c, -A.h, A.h); }
    sOut = s; tOut = t; bOut = b; cOut = c; fOut = f;
}

/* Detection, per step: is any pair of shapes of bodies (ia, ib) within the contact band (1e-3 of slack for round-off: emitPairRows
   applies the exact test)?  Also folds the pair's surface distance into the running minimum `clearance`. */
template <typename T, typename Model>
TGB_D bool pairNear(const Model& M, const T* sPose, int ia, int ib, T& clearance) {
    const T* pa = sPose + 12 * ia;
    const T* pb = sPose + 12 * ib;
    {
        /* bounding spheres about the body origins */
        const V3<T> d(pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]);
        const T reach = M.body(BF_BOUND, ia) + M.body(BF_BOUND, ib) + pairBreakingThreshold<T>() + T(1e-3);
        if (d.length2() > reach * reach) return false;
    }
    bool near = false;
    const int sa0 = M.bodyi(BI_SHAPE_BEGIN, ia), sa1 = sa0 + M.bodyi(BI_SHAPE_COUNT, ia);
    const int sb0 = M.bodyi(BI_SHAPE_BEGIN, ib), sb1 = sb0 + M.bodyi(BI_SHAPE_COUNT, ib);
    for (int sa = sa0; sa < sa1; ++sa) {
        CoreSegment<T> ga;
        if (!coreSegment<T>(M, pa, sa, ga)) continue;
        for (int sb = sb0; sb < sb1; ++sb) {
            CoreSegment<T> gb;
            if (!coreSegment<T>(M, pb, sb, gb)) continue;
            T s, t, b, c, f;
            closestSegmentParams<T>(ga, gb, s, t, b, c, f);
            const V3<T> dv = (ga.p + ga.d * s) - (gb.p + gb.d * t);
            const T dist = tsqrtNoBranch(dv.length2()) - ga.r - gb.r;
            clearance = dist < clearance ? dist : clearance;
            near = near || dist <= pairBreakingThreshold<T>() + T(1e-3);
        }
    }
    return near;
}

/* Emission (rare): the contacts of bodies (ia, ib) as rows base, base + 1, ... of the environment's pair rows; returns the number
   written.  Mirrors the oracle's generateBodyPairContacts expression by expression. */
template <typename T, typename Model>
__device__ __noinline__ int emitPairRows(const Model M, const T* sPose, int ia, int ib, T* rows, int base, int cap) {
    const T* pa = sPose + 12 * ia;
    const T* pb = sPose + 12 * ib;
    int count = 0;
    const int sa0 = M.bodyi(BI_SHAPE_BEGIN, ia), sa1 = sa0 + M.bodyi(BI_SHAPE_COUNT, ia);
    const int sb0 = M.bodyi(BI_SHAPE_BEGIN, ib), sb1 = sb0 + M.bodyi(BI_SHAPE_COUNT, ib);
    T mu = M.body(BF_FRICTION, ia) * M.body(BF_FRICTION, ib);
    mu = mu < T(-10) ? T(-10) : (mu > T(10) ? T(10) : mu);
    const T restitution = M.body(BF_RESTITUTION, ia) * M.body(BF_RESTITUTION, ib);
    for (int sa = sa0; sa < sa1; ++sa) {
        CoreSegment<T> ga;
        if (!coreSegment<T>(M, pa, sa, ga)) continue;
        for (int sb = sb0; sb < sb1; ++sb) {
            CoreSegment<T> gb;
            if (!coreSegment<T>(M, pb, sb, gb)) continue;
            T s0, t0, b, c, f;
            closestSegmentParams<T>(ga, gb, s0, t0, b, c, f);
            T sp[2] = {s0, s0}, tp[2] = {t0, t0};
            int np = 1;
            if (ga.h > T(0) && gb.h > T(0) && T(1) - b * b < pairParallel<T>()) {
                /* B's end points in A's axis parameter: s(t) = t b - c */
                const T e0 = -gb.h * b - c, e1 = gb.h * b - c;
                const T emin = e0 < e1 ? e0 : e1, emax = e0 < e1 ? e1 : e0;
                const T lo = -ga.h > emin ? -ga.h : emin, hi = ga.h < emax ? ga.h : emax;
                if (hi - lo > T(1e-9)) {
                    np = 2;
                    sp[0] = lo; tp[0] = clampT(b * lo + f, -gb.h, gb.h);
                    sp[1] = hi; tp[1] = clampT(b * hi + f, -gb.h, gb.h);
                }
            }
            for (int k = 0; k < np; ++k) {
                const V3<T> ca = ga.p + ga.d * sp[k], cb = gb.p + gb.d * tp[k];
                const V3<T> dv = ca - cb;
                const T d = tsqrt(dv.length2());
                const T dist = d - ga.r - gb.r;
                if (dist > pairBreakingThreshold<T>()) continue;
                if (count >= 4 || base + count >= cap) continue;
                V3<T> n;
                if (d > T(1e-9)) n = dv * (T(1) / d);
                else {
                    /* the axes meet: any direction perpendicular to both */
                    const V3<T> x = ga.d.cross(gb.d);
                    const T xl = x.length();
                    if (xl > T(1e-9)) n = x * (T(1) / xl);
                    else {
                        /* btPlaneSpace1 of A's axis */
                        const V3<T> a = ga.d;
                        if (tabs(a.z) > T(0.7071067811865475244008443621048490)) {
                            const T kk = T(1) / tsqrt(a.y * a.y + a.z * a.z);
                            n = V3<T>(T(0), -a.z * kk, a.y * kk);
                        } else {
                            const T kk = T(1) / tsqrt(a.x * a.x + a.y * a.y);
                            n = V3<T>(-a.y * kk, a.x * kk, T(0));
                        }
                    }
                }
                const V3<T> pA = ca - n * ga.r, pB = cb + n * gb.r; /* the contact points on the two surfaces */
                T* row = rows + (base + count) * TGB_NPR;
                row[PR_N + 0] = n.x; row[PR_N + 1] = n.y; row[PR_N + 2] = n.z;
                row[PR_RA + 0] = pA.x; row[PR_RA + 1] = pA.y; row[PR_RA + 2] = pA.z;
                row[PR_RB + 0] = pB.x; row[PR_RB + 1] = pB.y; row[PR_RB + 2] = pB.z;
                row[PR_LAMN] = dist; row[PR_LAMPUSH] = restitution;
                row[PR_MU] = mu; row[PR_IDS] = (T)(ia + 32 * ib);
                ++count;
            }
        }
    }
    return count;
}

/* One environment's contact solve when it has rows between two bodies.  Called by ALL lanes of the environment's group (lanes without
   a body, static bodies and bodies without rows included); groupMask names them.  ct / nc / anyPenOwn: this lane's own rows against the
   static world (ground, obstacles) as generated by the step kernel; pr / nPairRows: the pair rows as emitPairRows left them.
   x, Iinv, invMass, vel, w, extLin, extAng: this lane's body (zeros for a lane without a dynamic body). */
template <typename T, bool kFast, bool OBST>
__device__ __noinline__ ContactSolveOut<T> solveCoupled(T* ct, int rowStride, int nc, bool anyPenOwn, V3<T> gn, T muGround, T* pr, int nPairRows,
                                                         int lane, int groupLane0, unsigned groupMask, V3<T> x, M3<T> Iinv, T invMass, V3<T> vel,
                                                         V3<T> w, V3<T> extLin, V3<T> extAng, T dt, int iterations) {
    /* ---- row set-up (setupContactConstraint / setupFrictionConstraint with two bodies) ---- */
    for (int p = 0; p < nPairRows; ++p) {
        T* row = pr + p * TGB_NPR;
        const int ids = (int)row[PR_IDS];
        const int ia = ids & 31, ib = ids >> 5;
        const bool isA = lane == ia, isB = lane == ib;
        const int src = groupLane0 + (isA ? ib : ia);
        const V3<T> n(row[PR_N + 0], row[PR_N + 1], row[PR_N + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r = V3<T>(row[po + 0], row[po + 1], row[po + 2]) - x;
        const T dist = row[PR_LAMN], combinedRestitution = row[PR_LAMPUSH];
        const V3<T> rxn = r.cross(n);
        const V3<T> angN = Iinv * rxn;
        const T dN = invMass + n.dot(angN.cross(r));
        const V3<T> vp = vel + w.cross(r);
        const T vdn = n.dot(vel + extLin) + rxn.dot(w + extAng);
        const T o_dN = __shfl_sync(groupMask, dN, src), o_vdn = __shfl_sync(groupMask, vdn, src);
        const V3<T> o_vp(__shfl_sync(groupMask, vp.x, src), __shfl_sync(groupMask, vp.y, src), __shfl_sync(groupMask, vp.z, src));
        const T denom = (isB ? o_dN : dN) + (isB ? dN : o_dN);
        const T invDiagN = kFast ? trcp(denom) : T(1) / denom;
        const V3<T> v = (isB ? o_vp : vp) - (isB ? vp : o_vp);
        const T relVel = n.dot(v);
        T restitution = combinedRestitution * -relVel;
        if (restitution <= T(0)) restitution = T(0);
        const V3<T> lat = v - n * relVel;
        const T lat2 = lat.length2();
        V3<T> t;
        if (lat2 > (kFast ? T(1.1920928955078125e-7) : T(2.2204460492503131e-16))) t = lat * (kFast ? trsqrt(lat2) : T(1) / tsqrt(lat2));
        else if (tabs(n.z) > T(0.7071067811865475244008443621048490)) {
            const T kk = T(1) / tsqrt(n.y * n.y + n.z * n.z);
            t = V3<T>(T(0), -n.z * kk, n.y * kk);
        } else {
            const T kk = T(1) / tsqrt(n.x * n.x + n.y * n.y);
            t = V3<T>(-n.y * kk, n.x * kk, T(0));
        }
        const T velDotN = (isB ? o_vdn : vdn) - (isB ? vdn : o_vdn);
        T velocityError = restitution - velDotN;
        const T penetration = dist;
        T erp = T(0.8);
        if (penetration > T(-0.04)) erp = T(0.2);
        T positionalError = T(0);
        if (penetration > T(0)) velocityError -= tdiv(penetration, dt);
        else positionalError = tdiv(-penetration * erp, dt);
        const T penetrationImpulse = positionalError * invDiagN;
        const T velocityImpulse = velocityError * invDiagN;
        T rhsN, rhsPen;
        if (penetration > T(-0.04)) { rhsN = penetrationImpulse + velocityImpulse; rhsPen = T(0); }
        else { rhsN = velocityImpulse; rhsPen = penetrationImpulse; }
        const V3<T> rxt = r.cross(t);
        const V3<T> angT = Iinv * rxt;
        const T dT = invMass + t.dot(angT.cross(r));
        const T rT = t.dot(vel + extLin) + rxt.dot(w);
        const T o_dT = __shfl_sync(groupMask, dT, src), o_rT = __shfl_sync(groupMask, rT, src);
        const T denomT = (isB ? o_dT : dT) + (isB ? dT : o_dT);
        const T relT = (isB ? o_rT : rT) - (isB ? rT : o_rT);
        const T invDiagT = kFast ? trcp(denomT) : T(1) / denomT;
        const T rhsT = (T(0) - relT) * invDiagT;
        __syncwarp(groupMask); /* every lane has read the detection's scratch fields of this row */
        if (isA) {
            row[PR_RA + 0] = r.x; row[PR_RA + 1] = r.y; row[PR_RA + 2] = r.z;
            row[PR_T + 0] = t.x; row[PR_T + 1] = t.y; row[PR_T + 2] = t.z;
            row[PR_INVDIAGN] = invDiagN; row[PR_RHSN] = rhsN; row[PR_RHSPEN] = rhsPen; row[PR_LAMN] = T(0); row[PR_LAMPUSH] = T(0);
            row[PR_INVDIAGT] = invDiagT; row[PR_RHST] = rhsT; row[PR_LAMT] = T(0);
        }
        if (isB) { row[PR_RB + 0] = r.x; row[PR_RB + 1] = r.y; row[PR_RB + 2] = r.z; }
    }
    __syncwarp(groupMask);
    bool anyPen = __ballot_sync(groupMask, anyPenOwn) != 0u;
    for (int p = 0; p < nPairRows; ++p) anyPen = anyPen || pr[p * TGB_NPR + PR_RHSPEN] != T(0);

    V3<T> dLin(T(0), T(0), T(0)), dAng(T(0), T(0), T(0)), push(T(0), T(0), T(0)), turn(T(0), T(0), T(0));
#define CTG(slot, f) ct[(slot) * rowStride + (f)]
    /* one visit of a pair row's normal (or split-impulse) constraint: `lin`, `ang` are this lane's accumulated velocity change */
    auto pairNormal = [&](T* row, int rhsField, int lamField, V3<T>& lin, V3<T>& ang) {
        const T rhs = row[rhsField];
        if (rhsField == PR_RHSPEN && rhs == T(0)) return; /* group-uniform */
        const int ids = (int)row[PR_IDS];
        const int ia = ids & 31, ib = ids >> 5;
        const bool isA = lane == ia, isB = lane == ib;
        const int src = groupLane0 + (isA ? ib : ia);
        const V3<T> n(row[PR_N + 0], row[PR_N + 1], row[PR_N + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r(row[po + 0], row[po + 1], row[po + 2]);
        const V3<T> rxn = r.cross(n);
        const T dvOwn = n.dot(lin) + rxn.dot(ang);
        const T dvOther = __shfl_sync(groupMask, dvOwn, src);
        const T dv = (isB ? dvOther : dvOwn) - (isB ? dvOwn : dvOther);
        T delta = rhs;
        delta -= dv * row[PR_INVDIAGN];
        const T lam = row[lamField];
        const T sum = lam + delta;
        T newLam = sum;
        if (sum < T(0)) { delta = -lam; newLam = T(0); }
        __syncwarp(groupMask);
        if (isA) row[lamField] = newLam;
        if (isA) { lin += n * (invMass * delta); ang += (Iinv * rxn) * delta; }
        if (isB) { lin -= n * (invMass * delta); ang -= (Iinv * rxn) * delta; }
    };
    auto pairFriction = [&](T* row) {
        const T lamN = row[PR_LAMN];
        if (!(lamN > T(0))) return; /* group-uniform */
        const int ids = (int)row[PR_IDS];
        const int ia = ids & 31, ib = ids >> 5;
        const bool isA = lane == ia, isB = lane == ib;
        const int src = groupLane0 + (isA ? ib : ia);
        const V3<T> t(row[PR_T + 0], row[PR_T + 1], row[PR_T + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r(row[po + 0], row[po + 1], row[po + 2]);
        const V3<T> rxt = r.cross(t);
        const T dvOwn = t.dot(dLin) + rxt.dot(dAng);
        const T dvOther = __shfl_sync(groupMask, dvOwn, src);
        const T dv = (isB ? dvOther : dvOwn) - (isB ? dvOwn : dvOther);
        const T mu = row[PR_MU];
        const T lo = -mu * lamN, hi = mu * lamN;
        T delta = row[PR_RHST];
        delta -= dv * row[PR_INVDIAGT];
        const T lam = row[PR_LAMT];
        const T sum = lam + delta;
        T newLam = sum;
        if (sum < lo) { delta = lo - lam; newLam = lo; }
        else if (sum > hi) { delta = hi - lam; newLam = hi; }
        __syncwarp(groupMask);
        if (isA) row[PR_LAMT] = newLam;
        if (isA) { dLin += t * (invMass * delta); dAng += (Iinv * rxt) * delta; }
        if (isB) { dLin -= t * (invMass * delta); dAng -= (Iinv * rxt) * delta; }
    };

    /* ---- split-impulse iterations ---- */
    if (anyPen)
        for (int it = 0; it < iterations; ++it) {
            for (int c = 0; c < nc; ++c) {
                const T rhsPen = CTG(c, CT_RHSPEN);
                if (rhsPen == T(0)) continue;
                const V3<T> n = OBST ? V3<T>(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2)) : gn;
                const V3<T> rxn(CTG(c, CT_RXN + 0), CTG(c, CT_RXN + 1), CTG(c, CT_RXN + 2));
                const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
                T delta = rhsPen;
                const T dv = n.dot(push) + rxn.dot(turn);
                delta -= dv * CTG(c, CT_INVDIAGN);
                const T lam = CTG(c, CT_LAMPUSH);
                const T sum = lam + delta;
                if (sum < T(0)) { delta = -lam; CTG(c, CT_LAMPUSH) = T(0); }
                else CTG(c, CT_LAMPUSH) = sum;
                push += n * (invMass * delta);
                turn += angN * delta;
            }
            for (int p = 0; p < nPairRows; ++p) pairNormal(pr + p * TGB_NPR, PR_RHSPEN, PR_LAMPUSH, push, turn);
            __syncwarp(groupMask);
        }
    /* ---- velocity iterations ---- */
    for (int it = 0; it < iterations; ++it) {
        for (int c = 0; c < nc; ++c) {
            const V3<T> n = OBST ? V3<T>(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2)) : gn;
            const V3<T> rxn(CTG(c, CT_RXN + 0), CTG(c, CT_RXN + 1), CTG(c, CT_RXN + 2));
            const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
            T delta = CTG(c, CT_RHSN);
            const T dv = n.dot(dLin) + rxn.dot(dAng);
            delta -= dv * CTG(c, CT_INVDIAGN);
            const T lam = CTG(c, CT_LAMN);
            const T sum = lam + delta;
            if (sum < T(0)) { delta = -lam; CTG(c, CT_LAMN) = T(0); }
            else CTG(c, CT_LAMN) = sum;
            dLin += n * (invMass * delta);
            dAng += angN * delta;
        }
        for (int p = 0; p < nPairRows; ++p) pairNormal(pr + p * TGB_NPR, PR_RHSN, PR_LAMN, dLin, dAng);
        __syncwarp(groupMask);
        for (int c = 0; c < nc; ++c) {
            const T lamN = CTG(c, CT_LAMN);
            if (!(lamN > T(0))) continue;
            const V3<T> t(CTG(c, CT_T + 0), CTG(c, CT_T + 1), CTG(c, CT_T + 2));
            const V3<T> rxt(CTG(c, CT_RXT + 0), CTG(c, CT_RXT + 1), CTG(c, CT_RXT + 2));
            const V3<T> angT(CTG(c, CT_ANGT + 0), CTG(c, CT_ANGT + 1), CTG(c, CT_ANGT + 2));
            const T mu = OBST ? CTG(c, CT_MU) : muGround;
            const T lo = -mu * lamN, hi = mu * lamN;
            T delta = CTG(c, CT_RHST);
            const T dv = t.dot(dLin) + rxt.dot(dAng);
            delta -= dv * CTG(c, CT_INVDIAGT);
            const T lam = CTG(c, CT_LAMT);
            const T sum = lam + delta;
            if (sum < lo) { delta = lo - lam; CTG(c, CT_LAMT) = lo; }
            else if (sum > hi) { delta = hi - lam; CTG(c, CT_LAMT) = hi; }
            else CTG(c, CT_LAMT) = sum;
            dLin += t * (invMass * delta);
            dAng += angT * delta;
        }
        for (int p = 0; p < nPairRows; ++p) pairFriction(pr + p * TGB_NPR);
        __syncwarp(groupMask);
    }
#undef CTG
    ContactSolveOut<T> o;
    o.dLin = dLin; o.dAng = dAng; o.push = push; o.turn = turn;
    return o;
}

} // namespace tgb
