/*
 * tg_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT). See tg_oracle.h for the
 * reference lines each function follows and for the parity status of each piece.
 *
 * Everything is double precision and sequential, in the reference's own evaluation order
 * (bodies in creation order, cables in addPair order, applyImpulse one cable at a time).
 */
#include "tg_oracle.h"
#include "tg_odeint.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    V3 operator+(const V3& o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3& o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(double s) const { return V3(x * s, y * s, z * s); }
    V3 operator/(double s) const { double i = 1.0 / s; return V3(x * i, y * i, z * i); } /* btVector3::operator/ multiplies by 1/s */
    V3& operator+=(const V3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    V3& operator-=(const V3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    V3& operator*=(double s) { x *= s; y *= s; z *= s; return *this; }
    double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
    V3 cross(const V3& o) const { return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    double length2() const { return dot(*this); }
    double length() const { return std::sqrt(length2()); }
};

struct Q4 {
    double x, y, z, w;
    Q4() : x(0), y(0), z(0), w(1) {}
    Q4(double a, double b, double c, double d) : x(a), y(b), z(c), w(d) {}
};

/* btQuaternion operator*(q1, q2) */
static Q4 qmul(const Q4& a, const Q4& b) {
    return Q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
              a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}

struct M3 {
    double m[3][3];
    V3 row(int i) const { return V3(m[i][0], m[i][1], m[i][2]); }
    V3 col(int j) const { return V3(m[0][j], m[1][j], m[2][j]); }
    V3 operator*(const V3& v) const { return V3(row(0).dot(v), row(1).dot(v), row(2).dot(v)); }
    M3 operator*(const M3& o) const {
        M3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
        return r;
    }
    M3 transpose() const {
        M3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = m[j][i];
        return r;
    }
    /* btMatrix3x3::scaled: column j multiplied by s[j] */
    M3 scaled(const V3& s) const {
        M3 r;
        for (int i = 0; i < 3; ++i) { r.m[i][0] = m[i][0] * s.x; r.m[i][1] = m[i][1] * s.y; r.m[i][2] = m[i][2] * s.z; }
        return r;
    }
    /* btMatrix3x3::setRotation */
    static M3 fromQuat(const Q4& q) {
        double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        double s = 2.0 / d;
        double xs = q.x * s, ys = q.y * s, zs = q.z * s;
        double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
        double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
        double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
        M3 r;
        r.m[0][0] = 1.0 - (yy + zz); r.m[0][1] = xy - wz;         r.m[0][2] = xz + wy;
        r.m[1][0] = xy + wz;         r.m[1][1] = 1.0 - (xx + zz); r.m[1][2] = yz - wx;
        r.m[2][0] = xz - wy;         r.m[2][1] = yz + wx;         r.m[2][2] = 1.0 - (xx + yy);
        return r;
    }
    /* btMatrix3x3::getRotation */
    Q4 toQuat() const {
        double trace = m[0][0] + m[1][1] + m[2][2];
        double t[4];
        if (trace > 0.0) {
            double s = std::sqrt(trace + 1.0);
            t[3] = s * 0.5;
            s = 0.5 / s;
            t[0] = (m[2][1] - m[1][2]) * s;
            t[1] = (m[0][2] - m[2][0]) * s;
            t[2] = (m[1][0] - m[0][1]) * s;
        } else {
            int i = m[0][0] < m[1][1] ? (m[1][1] < m[2][2] ? 2 : 1) : (m[0][0] < m[2][2] ? 2 : 0);
            int j = (i + 1) % 3, k = (i + 2) % 3;
            double s = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
            t[i] = s * 0.5;
            s = 0.5 / s;
            t[3] = (m[k][j] - m[j][k]) * s;
            t[j] = (m[j][i] + m[i][j]) * s;
            t[k] = (m[k][i] + m[i][k]) * s;
        }
        return Q4(t[0], t[1], t[2], t[3]);
    }
};

/* v * M (row vector times matrix), btVector3 operator*(v, m) */
static V3 vecMat(const V3& v, const M3& m) { return V3(m.col(0).dot(v), m.col(1).dot(v), m.col(2).dot(v)); }

struct Shape {
    int kind; /* tgb_shape_kind */
    double radius, halfLen, margin;
    double halfZ; /* box: third half extent */
    V3 pos;
    M3 rot;
};

struct Body {
    double mass, invMass;
    V3 invInertiaLocal;
    V3 pos;
    M3 basis;
    V3 linvel, angvel;
    M3 invInertiaWorld;
    V3 gravityForce; /* btRigidBody::m_gravity = acceleration * mass */
    V3 totalForce, totalTorque;
    double friction, rollingFriction, restitution;
    int shapeBegin, shapeCount;
    bool gyro;
    void updateInertiaTensor() { invInertiaWorld = basis.scaled(invInertiaLocal) * basis.transpose(); }
};

struct Cable {
    int a, b;
    V3 la, lb;
    double k, c;
    double rest, prev;
    double velocity, damping;
    int actuator;
    bool backdrivable;
    double maxTens, tVel, minAL, minRL;
    double mRadius, mFriction, mInertia;
    /* tgBasicActuator */
    double preferred;
    /* tgKinematicActuator */
    double motorVel, desiredTorque;
};

/* tgBulletContactSpringCable (src/core/tgBulletContactSpringCable.cpp) with its sliding anchors
   (src/core/tgBulletSpringCableAnchor.cpp); see include/ntrtsim_b200.h "Contact cables" for the restatement's terms. */
struct SlideAnchor {
    int body;
    V3 local;  /* attachedRelativeOriginalPosition */
    V3 normal; /* contactNormal in the BODY frame (tgBulletSpringCableAnchor.cpp:38 USE_BASIS: basis^-1 * cn at creation, :56-57), never
                  updated afterwards (:40 SKIP_CONTACT_UPDATE); the world normal, from the cable into the body, turns with the body */
};
struct ContactCable {
    int a, b;
    V3 la, lb;
    double k, c, rest, prev, velocity, damping, thickness, resolution;
    double maxTens, tVel, minAL, minRL, preferred; /* the tgBasicActuator around it */
    std::vector<SlideAnchor> slides; /* between the two permanent anchors, in order */
};
/* one point of the "manifold" of (cable, body) */
struct CablePoint { int body; V3 pos; V3 normal; double dist; };

struct Contact {
    int body;
    long long key = -1; /* identity of the manifold point across steps (persistent manifold): see warmStartKey */
    int bodyB = -1; /* body pairs: the second body (the normal points from B to A, Bullet's m_normalWorldOnB); -1 = the static world */
    V3 n;        /* ground -> rod */
    V3 r;        /* contact point on the rod relative to its centre of mass */
    V3 rB;       /* body pairs: contact point on B relative to B's centre of mass */
    V3 rxnB, angNB, rxtB, angTB; /* B's side of the two rows */
    double dist;
    /* normal row */
    V3 rxn, angN;
    double invDiagN, rhsN, rhsPen, lamN, lamPush;
    /* friction row */
    V3 t, rxt, angT;
    double invDiagT, rhsT, lamT, mu;
    double restitution; /* combined restitution of the pair (body x ground, or body x obstacle) */
};

/* Bullet 2.82 btContactSolverInfo defaults (tgWorldBulletPhysicsImpl.cpp:134-146 leaves them alone). */
const double kErp = 0.2, kErp2 = 0.8, kSplitThreshold = -0.04, kTurnErp = 0.1, kMaxGyro = 100.0;
const double kBreakingThreshold = 0.02;            /* gContactBreakingThreshold */
const double kAngularMotionThreshold = 0.5 * 1.57079632679489661923; /* 0.5*SIMD_HALF_PI */

/* btTransformUtil::integrateTransform */
static void integrateTransform(const V3& pos, const M3& basis, const V3& linvel, const V3& angvel, double dt,
                               V3& posOut, M3& basisOut) {
    posOut = pos + linvel * dt;
    V3 axis;
    double fAngle = angvel.length();
    if (fAngle * dt > kAngularMotionThreshold) fAngle = kAngularMotionThreshold / dt;
    if (fAngle < 0.001) {
        axis = angvel * (0.5 * dt - (dt * dt * dt) * 0.020833333333 * fAngle * fAngle);
    } else {
        axis = angvel * (std::sin(0.5 * fAngle * dt) / fAngle);
    }
    Q4 dorn(axis.x, axis.y, axis.z, std::cos(fAngle * dt * 0.5));
    Q4 orn0 = basis.toQuat();
    Q4 p = qmul(dorn, orn0);
    double l = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
    p.x /= l; p.y /= l; p.z /= l; p.w /= l;
    basisOut = M3::fromQuat(p);
}

/* btPlaneSpace1, first vector only */
static V3 planeSpace1(const V3& n) {
    if (std::fabs(n.z) > 0.7071067811865475244008443621048490) {
        double a = n.y * n.y + n.z * n.z;
        double k = 1.0 / std::sqrt(a);
        return V3(0, -n.z * k, n.y * k);
    }
    double a = n.x * n.x + n.y * n.y;
    double k = 1.0 / std::sqrt(a);
    return V3(-n.y * k, n.x * k, 0);
}

} // namespace

struct tgo_world {
    tgb_world_cfg cfg;
    tgb_ground_desc ground;
    std::vector<Body> bodies;
    std::vector<Shape> shapes;
    std::vector<Cable> cables;
    std::vector<ContactCable> contactCables;
    int modelBody0 = 0; /* first body of the model added last (tgo_add_contact_cables indexes relative to it) */
    std::vector<Contact> contacts;
    int ctrlKind = TGB_CTRL_NONE;
    std::vector<double> ctrlParams;
    double time = 0.0;     /* tgSimView-style accumulated time */
    double ctrlTime = 0.0; /* EscapeController::m_totalTime; SerializedSpineControl::simTime */
    double spineUpdateTime = 0.0; /* SerializedSpineControl::updateTime */
    bool trackMaxTension = false;
    std::vector<double> maxTension; /* running maximum of the tension history (hist = true) */
    V3 episodeStartCom;
    int iterations = 10;
    int status = 0;
    /* Persistent manifold + warm starting (btPersistentManifold keeps a point's applied impulses from step to step;
       btSequentialImpulseConstraintSolver starts the next solve from them scaled by btContactSolverInfo::m_warmstartingFactor = 0.85,
       SOLVER_USE_WARMSTARTING is on by default).  A point keeps its identity through its slot: (shape, rim point) against the ground,
       (obstacle shape, shape, sample) against an obstacle, (shape, shape, point) between two bodies.  0 = every solve starts from zero. */
    double warmStart = 0.0; /* same default as the device library (TGB_DEFAULT_WARM_STARTING, see include/ntrtsim_b200.h for why) */
    struct Cached { double lamN, lamT; };
    std::vector<std::pair<long long, Cached> > manifoldCache; /* of the last step, keyed (key, body) packed */
    int pairCap = 6;            /* rows of body-pair contacts per environment (tgb_set_pair_contact_cap); 0 = bodies do not collide */
    int pairRows = 0;           /* body-pair rows of the last world step */
    double minClearance = 1e30; /* smallest surface distance of two bodies' shapes seen by any world step */

    /* ---- CPG + impedance controller (TGB_CTRL_CPG_IMPEDANCE) ---------------------------------------------- */
    struct Cpg {
        int nNodes = 0;
        std::vector<int> begin, target, pair;
        std::vector<double> impedance; /* [n][4] */
        double rConst = 20, dMin = 0, dMax = 5, descCom = 2, controlTime = 0.001, cableStep = 1e-4, heightLo = 1, heightHi = 25;
        int maxSteps = 200, scoreBody = 0;
        double freqParam = 0, radiusParam = 0;
        std::vector<double> pairs;       /* [nPairs][2] weight, phase */
        std::vector<double> x;           /* [n][3] phi, r, rDot */
        std::vector<double> nodeValue;   /* r cos(phi) */
        double updateTime = 0;
        std::vector<double> cableTime;   /* tgCPGActuatorControl::m_controlTime of every cable */
        int numSteps = 0;
        /* tgCPGCableControl (cable_control 1): tgPIDController gains and state, last commanded tension of every cable */
        int cableControl = 0;
        double kP = 0, kI = 0, kD = 0;
        std::vector<double> pidPrevError, pidIntError, commandedTension;
    } cpg;

    /* CPGNode::nodeEquation, CPGNode.cpp:90-101 */
    static double nodeEquation(double d, double c0, double c1, double dMin, double dMax) {
        if (d >= dMin && d <= dMax) return c1 * d + c0;
        return 0;
    }
    /* integrate_function::operator() -> CPGEquations::updateNodes -> CPGNode::updateDTs, CPGNode.cpp:71-88 */
    void cpgRhs(const std::vector<double>& x, std::vector<double>& dxdt) {
        const int n = cpg.nNodes;
        for (int i = 0; i < n; ++i) {
            const double phi = x[3 * i], r = x[3 * i + 1], rd = x[3 * i + 2];
            double phiDot = 2 * M_PI * nodeEquation(cpg.descCom, cpg.freqParam, cpg.freqParam, cpg.dMin, cpg.dMax);
            for (int k = cpg.begin[i]; k < cpg.begin[i + 1]; ++k) {
                const int j = cpg.target[k];
                const double w = cpg.pairs[2 * cpg.pair[k]], th = cpg.pairs[2 * cpg.pair[k] + 1];
                phiDot += w * x[3 * j + 1] * std::sin(x[3 * j] - phi - th);
            }
            const double rDD = cpg.rConst * (cpg.rConst / 4 * (nodeEquation(cpg.descCom, cpg.radiusParam, cpg.radiusParam, cpg.dMin, cpg.dMax) - r) - rd);
            dxdt[3 * i] = phiDot;
            dxdt[3 * i + 1] = rd;
            dxdt[3 * i + 2] = rDD;
        }
        ++cpg.numSteps;
    }
    /* CPGEquations::update, CPGEquations.cpp:231-267 */
    void cpgUpdate(double dt) {
        double stepSize = dt <= 0.1 ? dt : 0.1;
        cpg.numSteps = 0;
        struct Sys {
            tgo_world* w;
            void operator()(const std::vector<double>& x, std::vector<double>& dxdt, double) { w->cpgRhs(x, dxdt); }
        } sys = {this};
        struct Obs {
            void operator()(const std::vector<double>&, double) {}
        } obs;
        tg_odeint::integrate(sys, cpg.x, 0.0, dt, stepSize, obs);
        for (int i = 0; i < cpg.nNodes; ++i) cpg.nodeValue[i] = cpg.x[3 * i + 1] * std::cos(cpg.x[3 * i]); /* CPGNode::updateNodeValues */
        if (cpg.numSteps > cpg.maxSteps) status |= 4; /* "Inefficient CPG Parameters" */
    }
    /* tgPIDController::control(dt, setPoint, sensorData), tgPIDController.cpp:80-109; the controllable is a tgKinematicActuator,
       whose setControlInput stores the desired torque (tgKinematicActuator.cpp:227-230) */
    void pidControl(int i, double dt, double setPoint, double sensorData) {
        const double error = setPoint - sensorData;
        cpg.pidIntError[i] += (error + cpg.pidPrevError[i]) / 2.0 * dt;
        const double dError = (error - cpg.pidPrevError[i]) / dt;
        const double result = cpg.kP * error + cpg.kI * cpg.pidIntError[i] + cpg.kD * dError;
        cables[i].desiredTorque = result;
        cpg.pidPrevError[i] = error;
    }

    /* tgCPGActuatorControl::onStep (tgCPGActuatorControl.cpp:79-100) / tgCPGCableControl::onStep (tgCPGCableControl.cpp:63-98)
       for cable i */
    void cpgCableControl(int i, double dt) {
        Cable& c = cables[i];
        cpg.cableTime[i] += dt;
        const double* g = &cpg.impedance[4 * i];
        if (c.actuator == TGB_ACT_KINEMATIC) { /* tgCPGCableControl with usePID */
            if (cpg.cableTime[i] >= cpg.cableStep) {
                /* tgImpedanceController::controlTension(tgBasicController&, ...), tgImpedanceController.cpp:118-150 */
                const double actualLength = cableLength(c);
                const double vel = c.motorVel * c.mRadius; /* tgKinematicActuator::getVelocity, tgKinematicActuator.cpp:181-184 */
                const double setTension = std::max(0.0, g[0] + g[1] * (actualLength - g[3]) + g[2] * (vel - cpg.nodeValue[i]));
                const double currentTension = cableTension(c);
                pidControl(i, dt, setTension, currentTension);
                cpg.commandedTension[i] = setTension;
                cpg.cableTime[i] = 0;
            } else {
                const double currentTension = cableTension(c);
                pidControl(i, dt, cpg.commandedTension[i], currentTension);
            }
            return;
        }
        if (cpg.cableTime[i] >= cpg.cableStep) {
            /* tgCPGActuatorControl hands the accumulated control time to the controller, tgCPGCableControl the step's dt */
            const double ct = cpg.cableControl == 1 ? dt : cpg.cableTime[i];
            /* tgImpedanceController::controlTension(tgBasicActuator&, ...), tgImpedanceController.cpp:163-185 */
            const double actualLength = cableLength(c);
            const double vel = c.velocity;
            double setTension = g[0] + g[1] * (actualLength - g[3]) + g[2] * (vel - cpg.nodeValue[i]);
            setTension = std::max(0.0, setTension);
            /* static tgTensionController::control(sca, dt, setPoint), tgTensionController.cpp:90-114 */
            const double currentTension = cableTension(c);
            const double delta = setTension - currentTension;
            const double diff = delta / c.k;
            double newLength = c.rest - diff;
            newLength = newLength < 0.1 ? 0.1 : newLength;
            setControlInput(c, newLength, ct);
            cpg.commandedTension[i] = setTension;
            cpg.cableTime[i] = 0;
        } else {
            moveMotors(c, dt);
        }
    }

    /* ---- cables ---------------------------------------------------------------------------- */
    V3 anchorWorld(const Body& b, const V3& local) const { return b.basis * local + b.pos; }

    double cableLength(const Cable& c) const {
        V3 d = anchorWorld(bodies[c.b], c.lb) - anchorWorld(bodies[c.a], c.la);
        return d.length();
    }
    double cableTension(const Cable& c) const {
        double t = (cableLength(c) - c.rest) * c.k;
        return t < 0.0 ? 0.0 : t;
    }

    static void applyImpulse(Body& b, const V3& J, const V3& rel) {
        if (b.invMass != 0.0) {
            b.linvel += J * b.invMass;
            b.angvel += b.invInertiaWorld * rel.cross(J);
        }
    }

    /* tgBulletSpringCable::calculateAndApplyForce */
    void cableForce(Cable& c, double dt) {
        Body& A = bodies[c.a];
        Body& B = bodies[c.b];
        V3 p1 = anchorWorld(A, c.la), p2 = anchorWorld(B, c.lb);
        V3 dist = p2 - p1;
        double L = dist.length();
        V3 unit = dist / L;
        double stretch = L - c.rest;
        double magnitude = c.k * stretch;
        double deltaStretch = L - c.prev;
        c.velocity = deltaStretch / dt;
        c.damping = c.c * c.velocity;
        if (std::fabs(magnitude) * 1.0 < std::fabs(c.damping))
            c.damping = (c.damping > 0.0 ? magnitude * 1.0 : -magnitude * 1.0);
        magnitude += c.damping;
        V3 force(0, 0, 0);
        if (L > c.rest) force = unit * magnitude;
        c.prev = L;
        applyImpulse(A, force * dt, p1 - A.pos);
        applyImpulse(B, -force * dt, p2 - B.pos);
    }


    /* ---- contact cables ---------------------------------------------------------------------- */
    enum { kMaxSliding = TGB_CCABLE_MAX_SLIDING, kMaxCablePoints = 64 };
    V3 toLocal(const Body& b, const V3& world) const { return vecMat(world - b.pos, b.basis); } /* basis^T (world - pos) */
    /* tgBulletSpringCableAnchor::getContactNormal (:202-214): the stored normal turned into the world frame and normalised */
    V3 ccWorldNormal(const SlideAnchor& q) const {
        const V3 n = bodies[q.body].basis * q.normal;
        const double l = n.length();
        return l > 0.0 ? n / l : V3(0, 0, 0);
    }
    /* world positions of the anchor chain, tgBulletContactSpringCable::m_anchors */
    void ccChain(const ContactCable& cc, std::vector<V3>& P) const {
        P.clear();
        P.push_back(anchorWorld(bodies[cc.a], cc.la));
        for (const SlideAnchor& q : cc.slides) P.push_back(anchorWorld(bodies[q.body], q.local));
        P.push_back(anchorWorld(bodies[cc.b], cc.lb));
    }
    /* getActualLength, tgBulletContactSpringCable.cpp:84-95 */
    static double ccLength(const std::vector<V3>& P) {
        double L = 0.0;
        for (size_t i = 0; i + 1 < P.size(); ++i) L += (P[i] - P[i + 1]).length();
        return L;
    }
    double contactCableLength(const ContactCable& cc) const { std::vector<V3> P; ccChain(cc, P); return ccLength(P); }
    double contactCableTension(const ContactCable& cc) const {
        const double t = (contactCableLength(cc) - cc.rest) * cc.k;
        return t < 0.0 ? 0.0 : t;
    }
    /* tgBasicActuator::setControlInput(input, dt) -> moveMotors (tgBasicActuator.cpp:149-225) on a contact cable */
    int contactCableControlInput(ContactCable& c, double input, double dt) {
        if (input < 0.0) return TGB_ERR_INVALID_ARGUMENT;
        c.preferred = input;
        const double stepSize = c.tVel * dt;
        const double actual = contactCableLength(c);
        if ((actual - c.preferred) * c.k > c.maxTens) c.preferred = actual - c.maxTens / c.k;
        const double diff = c.preferred - c.rest;
        const double fabsDiff = std::fabs(diff);
        if ((actual > c.minAL) || (diff > 0)) {
            if (std::fabs(diff) > stepSize) c.rest += (diff / fabsDiff) * stepSize;
            else c.rest += diff;
        }
        c.rest = (c.rest > c.minRL) ? c.rest : c.minRL;
        return 0;
    }
    /* the manifold point of `body` nearest to p: its distance to p (INFINITY when the body has none) and its normal */
    static double ccNearest(const std::vector<CablePoint>& pts, int body, const V3& p, V3& normal, V3& pos) {
        double best2 = INFINITY;
        const CablePoint* bm = nullptr;
        for (const CablePoint& m : pts) {
            if (m.body != body) continue;
            const double d2 = (m.pos - p).length2();
            if (d2 < best2) { best2 = d2; bm = &m; }
        }
        if (!bm) return INFINITY;
        normal = bm->normal; pos = bm->pos;
        return std::sqrt(best2);
    }
    /* tgBulletSpringCableAnchor::getManifoldDistance (:279-334) of a sliding anchor: distance to the nearest point of its body's
       manifold, and that point's normal when it lies within 0.1 (else the anchor's own) */
    double ccManifoldDistance(const std::vector<CablePoint>& pts, const SlideAnchor& q, int body, V3& normalOut) const {
        normalOut = ccWorldNormal(q);
        if (q.body != body) return INFINITY;
        V3 n, pos;
        const double len = ccNearest(pts, body, anchorWorld(bodies[q.body], q.local), n, pos);
        if (len < 0.1) normalOut = n;
        return len;
    }
    /* tgBulletSpringCableAnchor::updateManifold (:232-277) when the manifold of (cable, body) is the one the anchor already holds:
       accepted unless the nearest point's normal opposes the anchor's */
    bool ccUpdateManifold(const std::vector<CablePoint>& pts, const SlideAnchor& q, int body) const {
        if (q.body != body) return false;
        V3 newNormal;
        ccManifoldDistance(pts, q, body, newNormal);
        return !((newNormal + ccWorldNormal(q)).length() < 0.5); /* the reference compares in the body frame: the same length */
    }
    /* the segment of the chain nearest to pos: the restatement's findNearestPastAnchor (:799-937) */
    static int ccNearestSegment(const std::vector<V3>& P, const V3& pos) {
        int best = 0;
        double bestD = INFINITY;
        for (size_t j = 0; j + 1 < P.size(); ++j) {
            const V3 e = P[j + 1] - P[j];
            const double e2 = e.length2();
            double t = e2 > 0.0 ? (pos - P[j]).dot(e) / e2 : 0.0;
            t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
            const double d = (P[j] + e * t - pos).length2();
            if (d < bestD) { bestD = d; best = (int)j; }
        }
        return best;
    }
    /* tgBulletSpringCableAnchor::setWorldPosition (:89-184) */
    bool ccSetWorldPosition(const std::vector<CablePoint>& pts, SlideAnchor& q, const V3& newPos) {
        const Body& b = bodies[q.body];
        const V3 current = anchorWorld(b, q.local);
        const V3 own = ccWorldNormal(q);
        V3 newNormal = own, contactPos = current;
        const double length = ccNearest(pts, q.body, newPos, newNormal, contactPos);
        if (length < 0.1) {
            q.local = toLocal(b, newPos);
            return !((newNormal + own).length() < 0.5);
        }
        return (current - contactPos).length() <= 0.1;
    }
    /* updateAnchorPositions (:484-561); returns the number of anchors removed */
    int ccUpdateAnchorPositions(ContactCable& cc, const std::vector<CablePoint>& pts) {
        int numPruned = 0;
        size_t i = 0;
        std::vector<V3> P;
        while (i < cc.slides.size()) {
            ccChain(cc, P);
            const V3 back = P[i], current = P[i + 1], forward = P[i + 2];
            const V3 lineA = forward - current, lineB = back - current;
            const V3 n = ccWorldNormal(cc.slides[i]);
            V3 tangentMove(0, 0, 0);
            const V3 along = lineB - lineA;
            const double al = along.length();
            if (al > 0.0) {
                const V3 td = (along / al).cross(n);
                const double tl = td.length();
                if (tl > 0.0) {
                    const V3 tangentDir = td / tl;
                    tangentMove = tangentDir * (lineB + lineA).dot(tangentDir);
                }
            }
            const V3 newPos = current + tangentMove;
            if (!ccSetWorldPosition(pts, cc.slides[i], newPos)) {
                cc.slides.erase(cc.slides.begin() + i);
                ++numPruned;
            } else ++i;
        }
        return numPruned;
    }
    /* the "manifolds" of a cable whose anchors stand at P: closest-point pairs of its segments with every shape of every body nearer
       than the contact breaking threshold (see include/ntrtsim_b200.h "Contact cables") */
    void ccDetect(const std::vector<V3>& P, double thickness, std::vector<CablePoint>& pts) const {
        pts.clear();
        for (size_t i = 0; i + 1 < P.size(); ++i) {
            const V3 e = P[i + 1] - P[i];
            const double len = e.length();
            Segment gs;
            gs.p = (P[i] + P[i + 1]) * 0.5;
            gs.d = len > 1e-12 ? e / len : V3(0, 1, 0);
            gs.h = len > 1e-12 ? 0.5 * len : 0.0;
            gs.r = thickness;
            for (size_t bi = 0; bi < bodies.size(); ++bi) {
                const Body& b = bodies[bi];
                for (int si = b.shapeBegin; si < b.shapeBegin + b.shapeCount; ++si) {
                    Segment gb;
                    if (!shapeSegment(b, shapes[si], gb)) continue;
                    double s0, t0, bb, c0, f0;
                    closestSegmentParams(gs, gb, s0, t0, bb, c0, f0);
                    const V3 ca = gs.p + gs.d * s0, cb = gb.p + gb.d * t0;
                    const V3 dv = ca - cb;
                    const double d = dv.length();
                    if (!(d > 1e-9)) continue; /* the axes meet: no direction */
                    const double dist = d - gb.r - gs.r;
                    if (dist > kBreakingThreshold) continue;
                    if ((int)pts.size() >= kMaxCablePoints) continue;
                    const V3 u = dv / d;
                    CablePoint m;
                    m.body = (int)bi; m.pos = cb + u * gb.r; m.normal = -u; m.dist = dist;
                    pts.push_back(m);
                }
            }
        }
    }
    /* tgBulletContactSpringCable::step (:97-125) */
    void contactCableStep(ContactCable& cc, double dt) {
        std::vector<V3> P;
        ccChain(cc, P);
        /* updateManifolds: this step's contact points of the segments with every body */
        std::vector<CablePoint> pts;
        ccDetect(P, cc.thickness, pts);
        /* updateManifolds + updateAnchorList: proposals of the penetrating points, one after the other */
        for (size_t pi = 0; pi < pts.size(); ++pi) {
            const CablePoint m = pts[pi];
            if (!(m.dist < 0.0)) continue;
            ccChain(cc, P);
            const int j = ccNearestSegment(P, m.pos); /* back anchor = chain index j, forward = j + 1 */
            const V3 lineA = P[j + 1] - m.pos, lineB = P[j] - m.pos;
            const double lengthA = lineA.length(), lengthB = lineB.length();
            SlideAnchor* backS = j >= 1 ? &cc.slides[j - 1] : nullptr;                          /* NULL: permanent */
            SlideAnchor* fwdS = j + 1 <= (int)cc.slides.size() ? &cc.slides[j] : nullptr;
            V3 tmp;
            const double mDistB = backS ? ccManifoldDistance(pts, *backS, m.body, tmp) : INFINITY;
            const double mDistA = fwdS ? ccManifoldDistance(pts, *fwdS, m.body, tmp) : INFINITY;
            const int backBody = backS ? backS->body : cc.a, fwdBody = fwdS ? fwdS->body : cc.b;
            bool del = false;
            if (lengthB <= cc.resolution && m.body == backBody && mDistB < mDistA)
                if (backS && ccUpdateManifold(pts, *backS, m.body)) del = true;
            if (lengthA <= cc.resolution && m.body == fwdBody && (!del || mDistA < mDistB))
                if (fwdS && ccUpdateManifold(pts, *fwdS, m.body)) del = true;
            if (del) continue;
            if (lineA.dot(m.normal) < 0.0 || lineB.dot(m.normal) < 0.0) continue;
            const V3 backNormal = backS ? ccWorldNormal(*backS) : V3(0, 0, 0), fwdNormal = fwdS ? ccWorldNormal(*fwdS) : V3(0, 0, 0);
            if ((backNormal.dot(m.normal) < 0.0 && m.body == backBody) || (fwdNormal.dot(m.normal) < 0.0 && m.body == fwdBody)) continue;
            if ((int)cc.slides.size() >= kMaxSliding) continue;
            SlideAnchor q;
            q.body = m.body; q.local = toLocal(bodies[m.body], m.pos); q.normal = vecMat(m.normal, bodies[m.body].basis); /* basis^-1 * cn */
            cc.slides.insert(cc.slides.begin() + j, q);
            if (contactCableLength(cc) > cc.prev + 2.0 * cc.resolution) cc.slides.erase(cc.slides.begin() + j);
        }
        /* pruneAnchors */
        int numPruned = 0, passes = 0;
        for (size_t i = 0; i < cc.slides.size();) {
            V3 n, pos;
            if (ccNearest(pts, cc.slides[i].body, anchorWorld(bodies[cc.slides[i].body], cc.slides[i].local), n, pos) == INFINITY) {
                cc.slides.erase(cc.slides.begin() + i);
                ++numPruned;
            } else ++i;
        }
        while (numPruned > 0 || passes <= 3) {
            numPruned = ccUpdateAnchorPositions(cc, pts);
            if (numPruned == 0) {
                size_t i = 0;
                while (i < cc.slides.size()) {
                    ccChain(cc, P);
                    const V3 lineA = P[i + 2] - P[i + 1], lineB = P[i] - P[i + 1];
                    const V3 n = ccWorldNormal(cc.slides[i]);
                    double nv1, nv2;
                    const double half = cc.resolution / 2.0;
                    if (lineA.length2() < half * half || lineB.length2() < half * half) nv1 = nv2 = -1.0;
                    else { nv1 = lineA.dot(n); nv2 = lineB.dot(n); }
                    if (nv1 < 0.0 || nv2 < 0.0) { cc.slides.erase(cc.slides.begin() + i); ++numPruned; }
                    else ++i;
                }
            }
            if (numPruned == 0) ++passes; else passes = 0;
        }
        /* calculateAndApplyForce (:127-194) */
        ccChain(cc, P);
        const double currLength = ccLength(P);
        double tension = (currLength - cc.rest) * cc.k;
        if (tension < 0.0) tension = 0.0;
        cc.velocity = (currLength - cc.prev) / dt;
        cc.damping = cc.c * cc.velocity;
        if (std::fabs(tension) * 1.0 < std::fabs(cc.damping)) cc.damping = (cc.damping > 0.0 ? tension * 1.0 : -tension * 1.0);
        const double magnitude = tension + cc.damping;
        const size_t n = P.size();
        for (size_t i = 0; i < n; ++i) {
            V3 force(0, 0, 0);
            if (i == 0) { const V3 d = P[1] - P[0]; const double l = d.length(); if (l > 0.0) force = (d / l) * magnitude; }
            else if (i == n - 1) { const V3 d = P[i] - P[i - 1]; const double l = d.length(); if (l > 0.0) force = -(d / l) * magnitude; }
            else {
                const V3 first = P[i + 1] - P[i], second = P[i - 1] - P[i];
                const double l1 = first.length(), l2 = second.length();
                V3 dir(0, 0, 0);
                if (l1 > 0.0) dir += first / l1;
                if (l2 > 0.0) dir += second / l2;
                force = dir * magnitude;
            }
            Body& body = bodies[i == 0 ? cc.a : (i == n - 1 ? cc.b : cc.slides[i - 1].body)];
            applyImpulse(body, force * dt, P[i] - body.pos);
        }
        cc.prev = currLength;
    }

    /* tgBasicActuator::moveMotors */
    void moveMotors(Cable& c, double dt) {
        double stepSize = c.tVel * dt;
        double actual = cableLength(c);
        if ((actual - c.preferred) * c.k > c.maxTens) c.preferred = actual - c.maxTens / c.k;
        double diff = c.preferred - c.rest;
        double fabsDiff = std::fabs(diff);
        if ((actual > c.minAL) || (diff > 0)) {
            if (std::fabs(diff) > stepSize) c.rest += (diff / fabsDiff) * stepSize;
            else c.rest += diff;
        }
        c.rest = (c.rest > c.minRL) ? c.rest : c.minRL;
    }
    /* tgBasicActuator::setControlInput(input, dt) */
    int setControlInput(Cable& c, double input, double dt) {
        if (input < 0.0) return TGB_ERR_INVALID_ARGUMENT;
        c.preferred = input;
        moveMotors(c, dt);
        return 0;
    }
    /* tgKinematicActuator::integrateRestLength + getAppliedTorque */
    void integrateRestLength(Cable& c, double dt) {
        double tension = cableTension(c);
        double maxTorque = c.maxTens * c.mRadius * (1.0 - c.mRadius * std::fabs(c.motorVel) / c.tVel);
        maxTorque = maxTorque < 0.0 ? 0.0 : maxTorque;
        double applied = std::fabs(c.desiredTorque) < maxTorque ? c.desiredTorque
                                                                 : c.desiredTorque / std::fabs(c.desiredTorque) * maxTorque;
        double acc = (applied - c.mFriction * c.motorVel + tension * c.mRadius) / c.mInertia;
        if (!c.backdrivable && acc * applied <= 0.0)
            c.motorVel = c.motorVel + acc * dt > 0.0 ? 0.0 : c.motorVel + acc * dt;
        else
            c.motorVel += acc * dt;
        c.rest += c.mRadius * c.motorVel * dt;
        c.rest = (c.rest > c.minRL) ? c.rest : c.minRL;
    }

    /* ---- controllers (model-level observers, run before the children step) ------------------- */
    void controllersStep(double dt) {
        const int C = (int)cables.size();
        switch (ctrlKind) {
        case TGB_CTRL_SETPOINT: /* tsTestRig.cpp:187 */
            for (int i = 0; i < C; ++i)
                if (cables[i].actuator == TGB_ACT_BASIC) setControlInput(cables[i], ctrlParams[i], dt);
            break;
        case TGB_CTRL_TORQUE:   /* tsTestRig.cpp:183 */
            for (int i = 0; i < C; ++i)
                if (cables[i].actuator == TGB_ACT_KINEMATIC) cables[i].desiredTorque = ctrlParams[i];
            break;
        case TGB_CTRL_TENSION: { /* T6TensionController.cpp:69-83 -> tgTensionController.cpp:50-88 */
            double setPoint = ctrlParams[0];
            for (int i = 0; i < C; ++i) {
                Cable& c = cables[i];
                if (c.actuator != TGB_ACT_BASIC) continue;
                double currentTension = cableTension(c);
                double delta = setPoint - currentTension;
                double diff = delta / c.k;
                double newLength = c.rest - diff;
                newLength = newLength < 0.0 ? 0.0 : newLength;
                setControlInput(c, newLength, dt);
            }
            break;
        }
        case TGB_CTRL_ESCAPE_SINE: { /* EscapeController.cpp:100-117,245-291 */
            ctrlTime += dt;
            const double* P = ctrlParams.data(); /* [8][4] amplitude, angFreq, phaseChange, dcOffset ; [32] = L0 */
            const double L0 = P[32];
            double phase = 0.0;
            int cluster = 0, oldCluster = 0;
            const int nMuscles = C < 24 ? C : 24;
            for (int i = 0; i < nMuscles; ++i) {
                oldCluster = cluster;
                if (i < 4) cluster = 0; else if (i < 8) cluster = 1; else if (i < 12) cluster = 2;
                else if (i < 16) cluster = 3; else if (i < 18) cluster = 4; else if (i < 20) cluster = 5;
                else if (i < 22) cluster = 6; else cluster = 7;
                double newLength = P[cluster * 4 + 0] * std::sin(P[cluster * 4 + 1] * ctrlTime + phase) + P[cluster * 4 + 3];
                double minLength = L0 * (1 - 0.50), maxLength = L0 * (1 + 0.50);
                if (newLength <= minLength) newLength = minLength;
                else if (newLength >= maxLength) newLength = maxLength;
                setControlInput(cables[i], newLength, dt);
                if (oldCluster != cluster) phase += P[cluster * 4 + 2];
            }
            for (int i = 0; i < C; ++i) moveMotors(cables[i], dt); /* EscapeController.cpp:108-116 */
            break;
        }
        case TGB_CTRL_SERIALIZED_SPINE: { /* SerializedSpineControl::onStep, SerializedSpineControl.cpp:262-371 */
            const double* P = ctrlParams.data();
            spineUpdateTime += dt;
            if (spineUpdateTime >= 1.0 / P[3]) {
                ctrlTime += spineUpdateTime;
                spineUpdateTime = 0.0;
            }
            for (int i = 0; i < C; ++i) {
                Cable& c = cables[i];
                if (c.actuator != TGB_ACT_BASIC) continue;
                const double* g = P + TGB_SPINE_HEADER + 5 * i;
                /* applyImpedanceControlGeneric, :174-205 */
                const double cycle = std::sin(ctrlTime * P[2] + g[4]);
                const double target = P[0] + cycle * P[1];
                /* tgImpedanceController::controlTension(tgBasicActuator&, ...), tgImpedanceController.cpp:163-185 */
                const double actualLength = cableLength(c);
                const double vel = c.velocity;
                const double setTension = std::max(0.0, g[0] + g[1] * (actualLength - g[3]) + g[2] * (vel - target));
                /* static tgTensionController::control(sca, dt, setPoint), tgTensionController.cpp:90-114 */
                const double currentTension = cableTension(c);
                const double delta = setTension - currentTension;
                const double diff = delta / c.k;
                double newLength = c.rest - diff;
                newLength = newLength < 0.1 ? 0.1 : newLength;
                setControlInput(c, newLength, dt);
            }
            /* the load, :336-370: btRigidBody::applyForce(force, marker relative position) accumulates total force and torque,
               which the next world step turns into external impulses and clears */
            const double t = ctrlTime;
            double scale = 0.0;
            if (t > P[4] && t < P[5]) scale = (t - P[4]) / (P[5] - P[4]);
            else if (t >= P[5] && t < P[6]) scale = 1.0;
            else if (t >= P[6] && t < P[7]) scale = (P[7] - t) / (P[7] - P[6]);
            const V3 force(P[8] * scale, P[9] * scale, P[10] * scale);
            const int nLoaded = (int)P[11];
            for (int k = 0; k < nLoaded; ++k) {
                const double* m = P + 12 + 5 * k;
                Body& b = bodies[(int)m[0]];
                const V3 f = force * m[4];
                const V3 rel = b.basis * V3(m[1], m[2], m[3]); /* abstractMarker::getRelativePosition */
                b.totalForce += f;
                b.totalTorque += rel.cross(f);
            }
            break;
        }
        case TGB_CTRL_CPG_IMPEDANCE: { /* BaseSpineCPGControl::onStep, BaseSpineCPGControl.cpp:253-279 */
            cpg.updateTime += dt;
            if (cpg.updateTime >= cpg.controlTime) {
                cpgUpdate(cpg.updateTime);
                cpg.updateTime = 0;
            }
            const double h = bodies[cpg.scoreBody].pos.y;
            if (h > cpg.heightHi || h < cpg.heightLo) status |= 2; /* bogus */
            break;
        }
        default: break;
        }
    }

    /* tgModel::step over the children: each actuator's step in addPair order */
    void modelsStep(double dt) {
        controllersStep(dt);
        for (size_t i = 0; i < cables.size(); ++i) {
            Cable& c = cables[i];
            if (ctrlKind == TGB_CTRL_CPG_IMPEDANCE && (c.actuator == TGB_ACT_BASIC || (c.actuator == TGB_ACT_KINEMATIC && cpg.cableControl == 1)))
                cpgCableControl((int)i, dt); /* notifyStep of the actuator */
            if (c.actuator == TGB_ACT_KINEMATIC) {
                integrateRestLength(c, dt);
                cableForce(c, dt);
                c.desiredTorque = 0.0; /* tgKinematicActuator.cpp:155-156 */
            } else {
                cableForce(c, dt);
            }
            /* logHistory (hist = true): the tension after the cable's step joins the history; only its maximum is kept */
            if (trackMaxTension) maxTension[i] = std::max(maxTension[i], cableTension(c));
        }
        /* tgBasicActuator::step of a contact cable = tgBulletContactSpringCable::step (no control input) */
        for (ContactCable& cc : contactCables) contactCableStep(cc, dt);
    }

    /* ---- world step (Bullet restatement) ----------------------------------------------------- */

    /* Analytic cylinder-vs-plane manifold: per cylinder end cap, four rim points of the margin-shrunk core
       cylinder, starting at the rim point deepest along -n; a point is a contact when its distance (after both
       margins) is within the contact breaking threshold; at most 4 per cylinder (btPersistentManifold cap). */
    /* Signed distance of a point to an oriented box (centre co, rotation Ro, half extents h) and the outward direction at the
       closest surface point: outside, the vector from the closest point; inside, the axis of the nearest face (lowest index on
       a tie). */
    static double boxDistance(const V3& p, const V3& co, const V3 e[3], const V3& h, V3& nOut) {
        const V3 d = p - co;
        const double q[3] = {e[0].dot(d), e[1].dot(d), e[2].dot(d)};
        const double hh[3] = {h.x, h.y, h.z};
        double a[3], o[3], len2 = 0.0;
        for (int i = 0; i < 3; ++i) {
            a[i] = std::fabs(q[i]) - hh[i];
            o[i] = a[i] > 0.0 ? a[i] : 0.0;
            len2 += o[i] * o[i];
        }
        double nl[3] = {0.0, 0.0, 0.0}, sd;
        if (len2 > 0.0) {
            sd = std::sqrt(len2);
            for (int i = 0; i < 3; ++i) nl[i] = (q[i] < 0.0 ? -o[i] : o[i]) / sd;
        } else {
            int k = 0;
            if (a[1] > a[k]) k = 1;
            if (a[2] > a[k]) k = 2;
            sd = a[k];
            nl[k] = q[k] < 0.0 ? -1.0 : 1.0;
        }
        nOut = e[0] * nl[0] + e[1] * nl[1] + e[2] * nl[2];
        return sd;
    }

    /* Static obstacles (tgCraterDeep, tgBlockField, ... are static tgBox bodies): sample points of every
       dynamic shape against every box shape of every static body.  Cylinder: five stations along the axis (both end caps, the
       quarter points and the middle), at each the surface point of the core cylinder that faces the box; sphere: its centre;
       box: its eight core corners.  A point is a contact when its distance (after the shape's margin; the obstacle counts with
       its full half extents) is within the contact breaking threshold; at most 4 per pair of shapes, in sampling order. */
    void generateObstacleContacts() {
        /* the device kernel keeps at most cap rows per body (tgb_capi.cu: 4 per shape of the largest dynamic body, times 1 + the
           number of obstacles up to two); the same cap here */
        int maxShapes = 1, nObst = 0;
        for (const Body& b : bodies) {
            if (b.invMass != 0.0) maxShapes = std::max(maxShapes, b.shapeCount);
            else
                for (int s = b.shapeBegin; s < b.shapeBegin + b.shapeCount; ++s) nObst += shapes[s].kind == TGB_SHAPE_BOX;
        }
        if (nObst == 0) return;
        const int cap = 4 * maxShapes * (1 + std::min(nObst, 2));
        std::vector<int> rows(bodies.size(), 0);
        for (const Contact& c : contacts) rows[c.body] += 1;
        for (size_t oi = 0; oi < bodies.size(); ++oi) {
            const Body& ob = bodies[oi];
            if (ob.invMass != 0.0) continue;
            for (int osi = ob.shapeBegin; osi < ob.shapeBegin + ob.shapeCount; ++osi) {
                const Shape& os = shapes[osi];
                if (os.kind != TGB_SHAPE_BOX) continue;
                const M3 Rm = ob.basis * os.rot;
                const V3 Ro[3] = {Rm.col(0), Rm.col(1), Rm.col(0).cross(Rm.col(1))};
                const V3 co = ob.basis * os.pos + ob.pos;
                const V3 h(os.radius, os.halfLen, os.halfZ);
                for (size_t bi = 0; bi < bodies.size(); ++bi) {
                    Body& b = bodies[bi];
                    if (b.invMass == 0.0) continue;
                    for (int si = b.shapeBegin; si < b.shapeBegin + b.shapeCount; ++si) {
                        const Shape& s = shapes[si];
                        const M3 R = b.basis * s.rot;
                        const V3 c = b.basis * s.pos + b.pos;
                        const V3 a = R.col(1);
                        const double rc = s.radius - s.margin, hc = s.halfLen - s.margin;
                        const int nPoints = s.kind == TGB_SHAPE_SPHERE ? 1 : s.kind == TGB_SHAPE_BOX ? 8 : 5;
                        int count = 0;
                        for (int k = 0; k < nPoints && count < 4 && rows[bi] < cap; ++k) {
                            V3 pc;
                            if (s.kind == TGB_SHAPE_SPHERE) {
                                pc = c;
                            } else if (s.kind == TGB_SHAPE_BOX) {
                                const V3 ex = R.col(0), ez = ex.cross(a);
                                const double hx = rc, hz = s.halfZ - s.margin;
                                pc = c + ex * ((k & 1) ? hx : -hx) + a * ((k & 2) ? hc : -hc) + ez * ((k & 4) ? hz : -hz);
                            } else {
                                const V3 pa = c + a * (hc * (0.5 * (double)(k - 2)));
                                V3 g;
                                boxDistance(pa, co, Ro, h, g);
                                const V3 radial = g - a * a.dot(g);
                                const double rl = radial.length();
                                pc = rl > 1e-9 ? pa - radial * (rc / rl) : pa;
                            }
                            V3 n;
                            const double dist = boxDistance(pc, co, Ro, h, n) - s.margin;
                            if (dist <= kBreakingThreshold) {
                                Contact ct;
                                ct.body = (int)bi;
                                ct.n = n;
                                ct.r = (pc - n * s.margin) - b.pos;
                                ct.dist = dist;
                                double mu = b.friction * ob.friction;
                                if (mu < -10.0) mu = -10.0;
                                if (mu > 10.0) mu = 10.0;
                                ct.mu = mu;
                                ct.restitution = b.restitution * ob.restitution;
                                ct.key = (((long long)1 << 40) | ((long long)osi << 24) | ((long long)si << 8) | (long long)k);
                                contacts.push_back(ct);
                                ++count;
                                rows[bi] += 1;
                            }
                        }
                    }
                }
            }
        }
    }

    void generateContacts() {
        contacts.clear();
        generateGroundContacts();
        generateObstacleContacts();
        /* rows are solved body by body in the device kernel: keep the same order here (stable: ground first, then obstacles) */
        std::stable_sort(contacts.begin(), contacts.end(), [](const Contact& x, const Contact& y) { return x.body < y.body; });
        /* rows between two bodies come last, in pair order: every iteration of the solver visits the rows of the static world
           first (body by body), then the pair rows */
        generateBodyPairContacts();
    }

    /* ---- contacts between bodies (tgWorldBulletPhysicsImpl.cpp:82-107: btAxisSweep3 + btDefaultCollisionConfiguration collide
       every pair of bodies of the world; tgRodInfo.cpp:99-113 btCylinderShape, tgSphereInfo.cpp:100-115 btSphereShape).
       Bullet runs GJK/EPA on the convex pair and keeps a persistent manifold; here, as for the ground, an analytic manifold: a
       cylinder counts as the capsule inscribed in it (its axis segment shortened by the radius at both ends, swept by the
       radius), a sphere as a segment of length zero.  One contact at the closest points of the two segments; two (at the ends of
       the overlap) when the segments are parallel within kParallel and overlap.  A point is a contact when the surface distance
       is within the contact breaking threshold.  At most 4 rows per pair of bodies and pairCap rows per environment, in the
       order (body A < body B, shape of A, shape of B, point).  Boxes take no part (static boxes are obstacles, see above). */
    struct Segment { V3 p, d; double h, r; };
    static bool shapeSegment(const Body& b, const Shape& s, Segment& g) {
        if (s.kind == TGB_SHAPE_BOX) return false;
        const M3 R = b.basis * s.rot;
        g.p = b.basis * s.pos + b.pos;
        g.d = R.col(1);
        g.r = s.radius;
        g.h = s.kind == TGB_SHAPE_SPHERE ? 0.0 : (s.halfLen - s.radius > 0.0 ? s.halfLen - s.radius : 0.0);
        return true;
    }
    static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
    /* closest points A(s) = pa + s da, B(t) = pb + t db of two segments given by centre, unit direction, half length */
    static void closestSegmentParams(const Segment& A, const Segment& B, double& sOut, double& tOut, double& bOut, double& cOut, double& fOut) {
        const V3 r = A.p - B.p;
        const double b = A.d.dot(B.d), c = A.d.dot(r), f = B.d.dot(r);
        const double denom = 1.0 - b * b;
        double s = 0.0, t = 0.0;
        if (A.h > 0.0 && denom > 1e-12) s = clampd((b * f - c) / denom, -A.h, A.h);
        t = b * s + f;
        if (t < -B.h) { t = -B.h; s = clampd(b * t - c, -A.h, A.h); }
        else if (t > B.h) { t = B.h; s = clampd(b * t - c, -A.h, A.h); }
        sOut = s; tOut = t; bOut = b; cOut = c; fOut = f;
    }
    void generateBodyPairContacts() {
        pairRows = 0; /* with pairCap == 0 the clearance is still measured, no row is generated */
        const double kParallel = 1e-4; /* 1 - cos^2 of the angle between the axes below which the overlap's two ends are used */
        const int nb = (int)bodies.size();
        for (int ia = 0; ia < nb; ++ia)
            for (int ib = ia + 1; ib < nb; ++ib) {
                Body& A = bodies[ia];
                Body& B = bodies[ib];
                if (A.invMass == 0.0 && B.invMass == 0.0) continue;
                int count = 0;
                for (int sa = A.shapeBegin; sa < A.shapeBegin + A.shapeCount; ++sa)
                    for (int sb = B.shapeBegin; sb < B.shapeBegin + B.shapeCount; ++sb) {
                        Segment ga, gb;
                        if (!shapeSegment(A, shapes[sa], ga) || !shapeSegment(B, shapes[sb], gb)) continue;
                        double s0, t0, b, c, f;
                        closestSegmentParams(ga, gb, s0, t0, b, c, f);
                        {
                            /* the surface distance of the pair = the distance at the closest points */
                            const V3 dv0 = (ga.p + ga.d * s0) - (gb.p + gb.d * t0);
                            const double clearance = std::sqrt(dv0.length2()) - ga.r - gb.r;
                            if (clearance < minClearance) minClearance = clearance;
                        }
                        double sp[2] = {s0, s0}, tp[2] = {t0, t0};
                        int np = 1;
                        if (ga.h > 0.0 && gb.h > 0.0 && 1.0 - b * b < kParallel) {
                            /* B's end points in A's axis parameter: s(t) = t b - c */
                            const double e0 = -gb.h * b - c, e1 = gb.h * b - c;
                            const double lo = std::max(-ga.h, std::min(e0, e1)), hi = std::min(ga.h, std::max(e0, e1));
                            if (hi - lo > 1e-9) {
                                np = 2;
                                sp[0] = lo; tp[0] = clampd(b * lo + f, -gb.h, gb.h);
                                sp[1] = hi; tp[1] = clampd(b * hi + f, -gb.h, gb.h);
                            }
                        }
                        for (int k = 0; k < np; ++k) {
                            const V3 ca = ga.p + ga.d * sp[k], cb = gb.p + gb.d * tp[k];
                            const V3 dv = ca - cb;
                            const double d2 = dv.length2();
                            const double d = std::sqrt(d2);
                            const double dist = d - ga.r - gb.r;
                            if (dist > kBreakingThreshold) continue;
                            if (count >= 4 || pairRows >= pairCap) continue;
                            V3 n;
                            if (d > 1e-9) n = dv * (1.0 / d);
                            else {
                                /* the axes meet: any direction perpendicular to both */
                                const V3 x = ga.d.cross(gb.d);
                                const double xl = x.length();
                                n = xl > 1e-9 ? x * (1.0 / xl) : planeSpace1(ga.d);
                            }
                            Contact ct;
                            ct.body = ia; ct.bodyB = ib;
                            ct.n = n;
                            ct.r = (ca - n * ga.r) - A.pos;
                            ct.rB = (cb + n * gb.r) - B.pos;
                            ct.dist = dist;
                            double mu = A.friction * B.friction;
                            if (mu < -10.0) mu = -10.0;
                            if (mu > 10.0) mu = 10.0;
                            ct.mu = mu;
                            ct.restitution = A.restitution * B.restitution;
                            ct.key = (((long long)2 << 40) | ((long long)sa << 24) | ((long long)sb << 8) | (long long)k);
                            contacts.push_back(ct);
                            ++count;
                            ++pairRows;
                        }
                    }
            }
    }

    void generateGroundContacts() {
        if (ground.kind == TGB_GROUND_EMPTY) return;
        const V3 n(ground.normal[0], ground.normal[1], ground.normal[2]);
        for (size_t bi = 0; bi < bodies.size(); ++bi) {
            Body& b = bodies[bi];
            if (b.invMass == 0.0) continue;
            for (int si = b.shapeBegin; si < b.shapeBegin + b.shapeCount; ++si) {
                const Shape& s = shapes[si];
                M3 R = b.basis * s.rot;
                V3 c = b.basis * s.pos + b.pos;
                V3 a = R.col(1);
                double rc = s.radius - s.margin, hc = s.halfLen - s.margin;
                double na = n.dot(a);
                V3 perp = -(n - a * na);
                double pl = perp.length();
                V3 u;
                if (pl > 1e-9) u = perp / pl;
                else u = R.col(0);
                V3 w = a.cross(u);
                int count = 0;
                if (s.kind == TGB_SHAPE_BOX) {
                    /* btBoxShape: the eight corners of the margin-shrunk core box, the face pair along local Y first ("end caps"),
                       around each face starting at the corner lowest along the normal; same distance test and cap of 4 */
                    const V3 ex = R.col(0), ez = ex.cross(a);
                    const double hx = s.radius - s.margin, hz = s.halfZ - s.margin;
                    const V3 bu = n.dot(ex) > 0.0 ? -ex : ex, bw = n.dot(ez) > 0.0 ? -ez : ez;
                    /* halves of the face diagonals: corner k = 0 is +d0 (lowest), 1 is +d1, 2 is -d0, 3 is -d1 */
                    const V3 d0 = bu * hx + bw * hz, d1 = bw * hz - bu * hx;
                    for (int e = 0; e < 2 && count < 4; ++e) {
                        V3 ce = c + a * (e == 0 ? -hc : hc);
                        for (int k = 0; k < 4 && count < 4; ++k) {
                            V3 pc = ce + ((k == 0) ? d0 : (k == 1) ? d1 : (k == 2) ? -d0 : -d1);
                            double dist = n.dot(pc) - ground.offset - s.margin;
                            if (dist <= kBreakingThreshold) {
                                Contact ct;
                                ct.body = (int)bi;
                                ct.n = n;
                                ct.r = (pc - n * s.margin) - b.pos;
                                ct.dist = dist;
                                double mu = b.friction * ground.friction;
                                if (mu < -10.0) mu = -10.0;
                                if (mu > 10.0) mu = 10.0;
                                ct.mu = mu;
                                ct.restitution = b.restitution * ground.restitution;
                                ct.key = ((long long)si << 8) | (long long)(e * 4 + k);
                                contacts.push_back(ct);
                                ++count;
                            }
                        }
                    }
                    continue;
                }
                /* btSphereShape = a point with margin radius (rc = hc = 0): one contact point */
                const int nEnds = s.kind == TGB_SHAPE_SPHERE ? 1 : 2, nDirs = s.kind == TGB_SHAPE_SPHERE ? 1 : 4;
                for (int e = 0; e < nEnds && count < 4; ++e) {
                    V3 ce = c + a * (e == 0 ? -hc : hc);
                    for (int k = 0; k < nDirs && count < 4; ++k) {
                        V3 dir = (k == 0) ? u : (k == 1) ? w : (k == 2) ? -u : -w;
                        V3 pc = ce + dir * rc;
                        double dist = n.dot(pc) - ground.offset - s.margin;
                        if (dist <= kBreakingThreshold) {
                            Contact ct;
                            ct.body = (int)bi;
                            ct.n = n;
                            ct.r = (pc - n * s.margin) - b.pos;
                            ct.dist = dist;
                            double mu = b.friction * ground.friction;
                            if (mu < -10.0) mu = -10.0;
                            if (mu > 10.0) mu = 10.0;
                            ct.mu = mu;
                            ct.restitution = b.restitution * ground.restitution;
                            ct.key = ((long long)si << 8) | (long long)(e * 4 + k);
                            contacts.push_back(ct);
                            ++count;
                        }
                    }
                }
            }
        }
    }

    void worldStep(double dt) {
        /* applyGravity */
        for (Body& b : bodies)
            if (b.invMass != 0.0) b.totalForce += b.gravityForce;

        /* performDiscreteCollisionDetection on the current transforms */
        generateContacts();

        /* solver bodies: external impulses (btSequentialImpulseConstraintSolver::initSolverBody + gyro) */
        const size_t nb = bodies.size();
        std::vector<V3> extLin(nb), extAng(nb), dLin(nb), dAng(nb), push(nb), turn(nb);
        for (size_t i = 0; i < nb; ++i) {
            Body& b = bodies[i];
            if (b.invMass == 0.0) continue;
            extLin[i] = b.totalForce * b.invMass * dt;
            extAng[i] = vecMat(b.totalTorque, b.invInertiaWorld) * dt;
            if (b.gyro) {
                V3 inertiaLocal(b.invInertiaLocal.x != 0 ? 1.0 / b.invInertiaLocal.x : 0,
                                b.invInertiaLocal.y != 0 ? 1.0 / b.invInertiaLocal.y : 0,
                                b.invInertiaLocal.z != 0 ? 1.0 / b.invInertiaLocal.z : 0);
                M3 Iw = b.basis.scaled(inertiaLocal) * b.basis.transpose();
                V3 tmp = Iw * b.angvel;
                V3 gf = b.angvel.cross(tmp);
                double l2 = gf.length2();
                if (l2 > kMaxGyro * kMaxGyro) gf *= 1.0 / std::sqrt(l2) * kMaxGyro;
                extAng[i] -= vecMat(gf, b.invInertiaWorld) * dt;
            }
        }

        /* contact rows (setupContactConstraint / setupFrictionConstraint, one friction direction) */
        for (Contact& c : contacts) {
            Body& b = bodies[c.body];
            const int i = c.body;
            c.rxn = c.r.cross(c.n);
            c.angN = b.invInertiaWorld * c.rxn;
            double denom = b.invMass + c.n.dot(c.angN.cross(c.r));
            V3 vel = b.linvel + b.angvel.cross(c.r);
            const bool pair = c.bodyB >= 0;
            const int j = pair ? c.bodyB : i;
            Body& bB = bodies[j];
            if (pair) {
                /* the second body of the row: Jacobian (-n, -(rB x n)); setupContactConstraint adds both bodies' terms */
                c.rxnB = c.rB.cross(c.n);
                c.angNB = bB.invInertiaWorld * c.rxnB;
                denom = denom + (bB.invMass + c.n.dot(c.angNB.cross(c.rB)));
                vel = vel - (bB.linvel + bB.angvel.cross(c.rB));
            }
            c.invDiagN = 1.0 / denom;
            double relVel = c.n.dot(vel);
            double combinedRestitution = c.restitution;
            double restitution = combinedRestitution * -relVel;
            if (restitution <= 0.0) restitution = 0.0;
            /* friction direction from the lateral velocity, else btPlaneSpace1 */
            V3 lat = vel - c.n * relVel;
            double lat2 = lat.length2();
            if (lat2 > DBL_EPSILON) c.t = lat * (1.0 / std::sqrt(lat2));
            else c.t = planeSpace1(c.n);
            c.lamN = 0.0; c.lamPush = 0.0; c.lamT = 0.0; /* warm start: applied below, once the row is complete */
            double velDotN = c.n.dot(b.linvel + extLin[i]) + c.rxn.dot(b.angvel + extAng[i]);
            if (pair) velDotN = velDotN - (c.n.dot(bB.linvel + extLin[j]) + c.rxnB.dot(bB.angvel + extAng[j]));
            double velocityError = restitution - velDotN;
            double penetration = c.dist;
            double erp = kErp2;
            if (penetration > kSplitThreshold) erp = kErp;
            double positionalError = 0.0;
            if (penetration > 0) velocityError -= penetration / dt;
            else positionalError = -penetration * erp / dt;
            double penetrationImpulse = positionalError * c.invDiagN;
            double velocityImpulse = velocityError * c.invDiagN;
            if (penetration > kSplitThreshold) { c.rhsN = penetrationImpulse + velocityImpulse; c.rhsPen = 0.0; }
            else { c.rhsN = velocityImpulse; c.rhsPen = penetrationImpulse; }
            /* friction row */
            c.rxt = c.r.cross(c.t);
            c.angT = b.invInertiaWorld * c.rxt;
            double denomT = b.invMass + c.t.dot(c.angT.cross(c.r));
            double relT = c.t.dot(b.linvel + extLin[i]) + c.rxt.dot(b.angvel);
            if (pair) {
                c.rxtB = c.rB.cross(c.t);
                c.angTB = bB.invInertiaWorld * c.rxtB;
                denomT = denomT + (bB.invMass + c.t.dot(c.angTB.cross(c.rB)));
                relT = relT - (c.t.dot(bB.linvel + extLin[j]) + c.rxtB.dot(bB.angvel));
            }
            c.invDiagT = 1.0 / denomT;
            c.rhsT = (0.0 - relT) * c.invDiagT;
            if (warmStart > 0.0) {
                /* setupContactConstraint / setFrictionConstraintImpulse: the point's impulses of the last step, scaled, are the starting
                   point of this solve and act on the bodies' velocity changes straight away (the friction impulse along the NEW direction) */
                const long long id = (c.key << 12) | ((long long)c.body << 6) | (long long)(c.bodyB < 0 ? 63 : c.bodyB);
                for (const auto& e : manifoldCache)
                    if (e.first == id) {
                        c.lamN = e.second.lamN * warmStart;
                        c.lamT = e.second.lamT * warmStart;
                        dLin[i] += c.n * (b.invMass * c.lamN) + c.t * (b.invMass * c.lamT);
                        dAng[i] += c.angN * c.lamN + c.angT * c.lamT;
                        if (pair) {
                            dLin[j] -= c.n * (bB.invMass * c.lamN) + c.t * (bB.invMass * c.lamT);
                            dAng[j] -= c.angNB * c.lamN + c.angTB * c.lamT;
                        }
                        break;
                    }
            }
        }

        /* split-impulse iterations, then velocity iterations (fixed-iteration PGS in place of Dantzig) */
        for (int it = 0; it < iterations; ++it) {
            for (Contact& c : contacts) {
                if (c.rhsPen == 0.0) continue;
                Body& b = bodies[c.body];
                const int i = c.body;
                double delta = c.rhsPen;
                double dv = c.n.dot(push[i]) + c.rxn.dot(turn[i]);
                if (c.bodyB >= 0) dv = dv - (c.n.dot(push[c.bodyB]) + c.rxnB.dot(turn[c.bodyB]));
                delta -= dv * c.invDiagN;
                double sum = c.lamPush + delta;
                if (sum < 0.0) { delta = -c.lamPush; c.lamPush = 0.0; }
                else c.lamPush = sum;
                push[i] += c.n * (b.invMass * delta);
                turn[i] += c.angN * delta;
                if (c.bodyB >= 0) {
                    push[c.bodyB] -= c.n * (bodies[c.bodyB].invMass * delta);
                    turn[c.bodyB] -= c.angNB * delta;
                }
            }
        }
        for (int it = 0; it < iterations; ++it) {
            for (Contact& c : contacts) {
                Body& b = bodies[c.body];
                const int i = c.body;
                double delta = c.rhsN;
                double dv = c.n.dot(dLin[i]) + c.rxn.dot(dAng[i]);
                if (c.bodyB >= 0) dv = dv - (c.n.dot(dLin[c.bodyB]) + c.rxnB.dot(dAng[c.bodyB]));
                delta -= dv * c.invDiagN;
                double sum = c.lamN + delta;
                if (sum < 0.0) { delta = -c.lamN; c.lamN = 0.0; }
                else c.lamN = sum;
                dLin[i] += c.n * (b.invMass * delta);
                dAng[i] += c.angN * delta;
                if (c.bodyB >= 0) {
                    dLin[c.bodyB] -= c.n * (bodies[c.bodyB].invMass * delta);
                    dAng[c.bodyB] -= c.angNB * delta;
                }
            }
            for (Contact& c : contacts) {
                if (!(c.lamN > 0.0)) continue;
                Body& b = bodies[c.body];
                const int i = c.body;
                double lo = -c.mu * c.lamN, hi = c.mu * c.lamN;
                double delta = c.rhsT;
                double dv = c.t.dot(dLin[i]) + c.rxt.dot(dAng[i]);
                if (c.bodyB >= 0) dv = dv - (c.t.dot(dLin[c.bodyB]) + c.rxtB.dot(dAng[c.bodyB]));
                delta -= dv * c.invDiagT;
                double sum = c.lamT + delta;
                if (sum < lo) { delta = lo - c.lamT; c.lamT = lo; }
                else if (sum > hi) { delta = hi - c.lamT; c.lamT = hi; }
                else c.lamT = sum;
                dLin[i] += c.t * (b.invMass * delta);
                dAng[i] += c.angT * delta;
                if (c.bodyB >= 0) {
                    dLin[c.bodyB] -= c.t * (bodies[c.bodyB].invMass * delta);
                    dAng[c.bodyB] -= c.angTB * delta;
                }
            }
        }

        if (warmStart > 0.0) {
            manifoldCache.clear();
            for (const Contact& c : contacts) {
                const long long id = (c.key << 12) | ((long long)c.body << 6) | (long long)(c.bodyB < 0 ? 63 : c.bodyB);
                manifoldCache.push_back(std::make_pair(id, Cached{c.lamN, c.lamT}));
            }
        }
        /* write back + split-impulse transform correction */
        for (size_t i = 0; i < nb; ++i) {
            Body& b = bodies[i];
            if (b.invMass == 0.0) continue;
            b.linvel = (b.linvel + dLin[i]) + extLin[i];
            b.angvel = (b.angvel + dAng[i]) + extAng[i];
            const V3& p = push[i];
            const V3& t = turn[i];
            if (p.x != 0 || p.y != 0 || p.z != 0 || t.x != 0 || t.y != 0 || t.z != 0) {
                V3 np; M3 nb2;
                integrateTransform(b.pos, b.basis, p, t * kTurnErp, dt, np, nb2);
                b.pos = np; b.basis = nb2;
            }
        }

        /* integrateTransforms: predictIntegratedTransform + proceedToTransform */
        for (Body& b : bodies) {
            if (b.invMass == 0.0) continue;
            V3 np; M3 nb2;
            integrateTransform(b.pos, b.basis, b.linvel, b.angvel, dt, np, nb2);
            b.pos = np; b.basis = nb2;
            b.updateInertiaTensor();
        }
        /* clearForces */
        for (Body& b : bodies) { b.totalForce = V3(); b.totalTorque = V3(); }
    }

    /* the point whose XZ displacement is the score: the scored segment's body under the CPG controller, else the
       mass-weighted centre of all rods (EscapeModel::getBallCOM) */
    V3 scoreCom() const {
        if (ctrlKind == TGB_CTRL_CPG_IMPEDANCE) return bodies[cpg.scoreBody].pos;
        return centerOfMass();
    }
    V3 centerOfMass() const {
        V3 s; double m = 0;
        for (const Body& b : bodies) { s += b.pos * b.mass; m += b.mass; }
        return m > 0 ? s / m : V3();
    }
};

static int g_defaultPairCap = 6;

extern "C" {

void tgo_set_default_pair_contact_cap(int cap) { g_defaultPairCap = cap < 0 ? 0 : cap; }

tgo_world* tgo_create(const tgb_world_cfg* cfg, const tgb_ground_desc* ground) {
    tgo_world* w = new tgo_world();
    w->pairCap = g_defaultPairCap;
    w->cfg = *cfg;
    w->ground = *ground;
    w->iterations = cfg->solver_iterations > 0 ? cfg->solver_iterations : 10;
    return w;
}

void tgo_destroy(tgo_world* w) { delete w; }

int tgo_add_model_flat(tgo_world* w, const tgb_body_desc* bodies, int n_bodies, const tgb_shape_desc* shapes,
                       int n_shapes, const tgb_cable_desc* cables, int n_cables) {
    const int body0 = (int)w->bodies.size();
    const int shape0 = (int)w->shapes.size();
    w->modelBody0 = body0;
    for (int i = 0; i < n_shapes; ++i) {
        Shape s;
        s.kind = shapes[i].kind;
        s.radius = shapes[i].radius; s.halfLen = shapes[i].half_length; s.margin = shapes[i].margin;
        s.halfZ = shapes[i].half_extent_z;
        s.pos = V3(shapes[i].pos[0], shapes[i].pos[1], shapes[i].pos[2]);
        s.rot = M3::fromQuat(Q4(shapes[i].quat[0], shapes[i].quat[1], shapes[i].quat[2], shapes[i].quat[3]));
        w->shapes.push_back(s);
    }
    for (int i = 0; i < n_bodies; ++i) {
        const tgb_body_desc& d = bodies[i];
        Body b;
        b.mass = d.mass;
        b.invMass = d.mass != 0.0 ? 1.0 / d.mass : 0.0; /* btRigidBody::setMassProps */
        b.invInertiaLocal = V3(d.inertia[0] != 0.0 ? 1.0 / d.inertia[0] : 0.0, d.inertia[1] != 0.0 ? 1.0 / d.inertia[1] : 0.0,
                               d.inertia[2] != 0.0 ? 1.0 / d.inertia[2] : 0.0);
        if (d.mass == 0.0) b.invInertiaLocal = V3();
        b.pos = V3(d.pos[0], d.pos[1], d.pos[2]);
        b.basis = M3::fromQuat(Q4(d.quat[0], d.quat[1], d.quat[2], d.quat[3]));
        b.friction = d.friction; b.rollingFriction = d.rolling_friction; b.restitution = d.restitution;
        b.shapeBegin = shape0 + d.shape_begin; b.shapeCount = d.shape_count;
        b.gyro = d.gyroscopic != 0;
        /* btRigidBody::setGravity: m_gravity = acceleration * (1 / m_inverseMass) */
        if (b.invMass != 0.0) b.gravityForce = V3(0, -w->cfg.gravity, 0) * (1.0 / b.invMass);
        b.updateInertiaTensor();
        w->bodies.push_back(b);
    }
    for (int i = 0; i < n_cables; ++i) {
        const tgb_cable_desc& d = cables[i];
        Cable c;
        c.a = body0 + d.body_a; c.b = body0 + d.body_b;
        c.la = V3(d.local_a[0], d.local_a[1], d.local_a[2]);
        c.lb = V3(d.local_b[0], d.local_b[1], d.local_b[2]);
        c.k = d.stiffness; c.c = d.damping;
        c.rest = d.rest_length; c.prev = d.prev_length;
        c.velocity = 0; c.damping = 0;
        c.actuator = d.actuator; c.backdrivable = d.backdrivable != 0;
        c.maxTens = d.max_tension; c.tVel = d.target_velocity; c.minAL = d.min_actual_length; c.minRL = d.min_rest_length;
        c.mRadius = d.motor_radius; c.mFriction = d.motor_friction; c.mInertia = d.motor_inertia;
        c.preferred = c.rest; /* tgBasicActuator ctor: m_preferredLength(m_restLength) */
        c.motorVel = 0; c.desiredTorque = 0;
        w->cables.push_back(c);
    }
    w->episodeStartCom = w->centerOfMass();
    return 0;
}

int tgo_add_contact_cables(tgo_world* w, const tgb_contact_cable_desc* cables, int n) {
    for (int i = 0; i < n; ++i) {
        const tgb_contact_cable_desc& d = cables[i];
        if (d.body_a < 0 || d.body_b < 0 || w->modelBody0 + d.body_a >= (int)w->bodies.size() || w->modelBody0 + d.body_b >= (int)w->bodies.size() ||
            !(d.stiffness > 0.0) || d.damping < 0.0 || d.rest_length < 0.0 || d.thickness < 0.0 || !(d.resolution > 0.0))
            return TGB_ERR_INVALID_ARGUMENT; /* tgBulletContactSpringCable::invariant, :939-952 */
        ContactCable cc;
        cc.a = w->modelBody0 + d.body_a; cc.b = w->modelBody0 + d.body_b;
        cc.la = V3(d.local_a[0], d.local_a[1], d.local_a[2]);
        cc.lb = V3(d.local_b[0], d.local_b[1], d.local_b[2]);
        cc.k = d.stiffness; cc.c = d.damping; cc.rest = d.rest_length; cc.prev = d.prev_length;
        cc.velocity = 0; cc.damping = 0; cc.thickness = d.thickness; cc.resolution = d.resolution;
        cc.maxTens = d.max_tension; cc.tVel = d.target_velocity; cc.minAL = d.min_actual_length; cc.minRL = d.min_rest_length;
        cc.preferred = cc.rest; /* tgBasicActuator ctor */
        w->contactCables.push_back(cc);
    }
    return 0;
}
int tgo_set_contact_cable_control_input(tgo_world* w, const double* input, double dt) {
    for (size_t i = 0; i < w->contactCables.size(); ++i) {
        if (input[i] != input[i]) continue; /* NaN: leave alone */
        const int rc = w->contactCableControlInput(w->contactCables[i], input[i], dt);
        if (rc != 0) return rc;
    }
    return 0;
}
/* the contact points the oracle's geometric manifold gives cable `cable` if its anchors stood at chain[n][3] (the reference-source
   harness feeds them to the reference's own anchor logic, oracle/ref): out [<= 64][8] = body, position on the body, normal from the
   cable into the body, distance; returns their number */
int tgo_contact_cable_points(const tgo_world* w, int cable, const double* chain, int n, double* out) {
    std::vector<V3> P;
    for (int i = 0; i < n; ++i) P.push_back(V3(chain[3 * i], chain[3 * i + 1], chain[3 * i + 2]));
    std::vector<CablePoint> pts;
    w->ccDetect(P, w->contactCables[cable].thickness, pts);
    for (size_t i = 0; i < pts.size(); ++i) {
        double* o = out + 8 * i;
        o[0] = pts[i].body; o[1] = pts[i].pos.x; o[2] = pts[i].pos.y; o[3] = pts[i].pos.z;
        o[4] = pts[i].normal.x; o[5] = pts[i].normal.y; o[6] = pts[i].normal.z; o[7] = pts[i].dist;
    }
    return (int)pts.size();
}
int tgo_num_contact_cables(const tgo_world* w) { return (int)w->contactCables.size(); }
void tgo_get_contact_cable_obs(const tgo_world* w, double* o) {
    for (size_t i = 0; i < w->contactCables.size(); ++i) {
        const ContactCable& cc = w->contactCables[i];
        o[5 * i + 0] = cc.rest;
        o[5 * i + 1] = w->contactCableLength(cc);
        o[5 * i + 2] = w->contactCableTension(cc);
        o[5 * i + 3] = cc.velocity;
        o[5 * i + 4] = (double)cc.slides.size();
    }
}
void tgo_get_contact_cable_anchors(const tgo_world* w, int cable, double* out) {
    const ContactCable& cc = w->contactCables[cable];
    for (int k = 0; k < TGB_CCABLE_MAX_SLIDING; ++k) {
        double* o = out + 7 * k;
        for (int j = 0; j < 7; ++j) o[j] = 0.0;
        o[0] = -1.0;
        if (k >= (int)cc.slides.size()) continue;
        const SlideAnchor& q = cc.slides[k];
        const V3 p = w->anchorWorld(w->bodies[q.body], q.local);
        const V3 wn = w->ccWorldNormal(q);
        o[0] = q.body; o[1] = p.x; o[2] = p.y; o[3] = p.z; o[4] = wn.x; o[5] = wn.y; o[6] = wn.z;
    }
}

int tgo_num_bodies(const tgo_world* w) { return (int)w->bodies.size(); }
int tgo_num_cables(const tgo_world* w) { return (int)w->cables.size(); }

void tgo_set_rod_state(tgo_world* w, const double* s) {
    /* (the persistent manifolds survive: the hybrid reference rollouts hand the state over every step) */
    for (size_t i = 0; i < w->bodies.size(); ++i, s += 13) {
        Body& b = w->bodies[i];
        b.pos = V3(s[0], s[1], s[2]);
        b.basis = M3::fromQuat(Q4(s[3], s[4], s[5], s[6]));
        b.linvel = V3(s[7], s[8], s[9]);
        b.angvel = V3(s[10], s[11], s[12]);
        b.updateInertiaTensor();
    }
    w->episodeStartCom = w->scoreCom();
}

void tgo_get_rod_state(const tgo_world* w, double* s) {
    for (size_t i = 0; i < w->bodies.size(); ++i, s += 13) {
        const Body& b = w->bodies[i];
        Q4 q = b.basis.toQuat();
        s[0] = b.pos.x; s[1] = b.pos.y; s[2] = b.pos.z;
        s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
        s[7] = b.linvel.x; s[8] = b.linvel.y; s[9] = b.linvel.z;
        s[10] = b.angvel.x; s[11] = b.angvel.y; s[12] = b.angvel.z;
    }
}

void tgo_set_cable_state(tgo_world* w, const double* rp) {
    for (size_t i = 0; i < w->cables.size(); ++i) {
        w->cables[i].rest = rp[2 * i];
        w->cables[i].prev = rp[2 * i + 1];
        w->cables[i].preferred = w->cables[i].rest;
    }
}

void tgo_get_cable_obs(const tgo_world* w, double* o) {
    for (size_t i = 0; i < w->cables.size(); ++i) {
        const Cable& c = w->cables[i];
        o[4 * i + 0] = c.rest;
        o[4 * i + 1] = w->cableLength(c);
        o[4 * i + 2] = w->cableTension(c);
        o[4 * i + 3] = c.actuator == TGB_ACT_KINEMATIC ? c.motorVel : c.velocity;
    }
}

int tgo_set_controller(tgo_world* w, int kind, const double* params, int n_params) {
    if (kind == TGB_CTRL_SERIALIZED_SPINE && n_params != TGB_SPINE_HEADER + 5 * (int)w->cables.size()) return TGB_ERR_INVALID_ARGUMENT;
    w->ctrlKind = kind;
    w->ctrlParams.assign(params, params + n_params);
    w->ctrlTime = 0.0;
    w->spineUpdateTime = 0.0;
    return 0;
}
void tgo_track_max_tension(tgo_world* w, int enable) {
    w->trackMaxTension = enable != 0;
    w->maxTension.assign(w->cables.size(), 0.0);
}
void tgo_get_max_tension(const tgo_world* w, double* out) {
    for (size_t i = 0; i < w->cables.size(); ++i) out[i] = i < w->maxTension.size() ? w->maxTension[i] : 0.0;
}

int tgo_set_control_input(tgo_world* w, const double* input, double dt) {
    for (size_t i = 0; i < w->cables.size(); ++i) {
        if (w->cables[i].actuator != TGB_ACT_BASIC) continue;
        int rc = w->setControlInput(w->cables[i], input[i], dt);
        if (rc) return rc;
    }
    return 0;
}

void tgo_world_step(tgo_world* w, double dt) { w->worldStep(dt); }
void tgo_models_step(tgo_world* w, double dt) { w->modelsStep(dt); }

int tgo_step(tgo_world* w, double dt, int n_steps) {
    if (!(dt > 0.0)) return TGB_ERR_INVALID_ARGUMENT;
    for (int s = 0; s < n_steps; ++s) {
        w->worldStep(dt);   /* tgSimulation.cpp:205 */
        w->modelsStep(dt);  /* tgSimulation.cpp:208-211 */
        w->time += dt;
    }
    return 0;
}

double tgo_get_time(const tgo_world* w) { return w->time; }
void tgo_mark_episode_start(tgo_world* w) { w->episodeStartCom = w->scoreCom(); w->status = 0; }
void tgo_get_scores(const tgo_world* w, double* s) {
    V3 c = w->scoreCom();
    double dx = c.x - w->episodeStartCom.x, dz = c.z - w->episodeStartCom.z;
    s[0] = std::sqrt(dx * dx + dz * dz);
    if (w->ctrlKind == TGB_CTRL_CPG_IMPEDANCE && (w->status & 2)) s[0] = -1.0; /* bogus, BaseSpineCPGControl.cpp:296-299 */
    s[1] = 0.0;
}
int tgo_get_status(const tgo_world* w) { return w->status; }
/* one CPGEquations::update(descCom, dt) on its own (known-answer test of the reference suite) */
void tgo_cpg_update(tgo_world* w, double dt) { w->cpgUpdate(dt); }
void tgo_get_cpg_state(const tgo_world* w, double* x) {
    for (size_t i = 0; i < w->cpg.x.size(); ++i) x[i] = w->cpg.x[i];
}
int tgo_set_cpg_controller(tgo_world* w, const tgb_cpg_desc* d, const double* params, int n_params) {
    if (!d || !params || n_params != 2 + 2 * d->n_pairs || d->n_nodes != (int)w->cables.size()) return TGB_ERR_INVALID_ARGUMENT;
    tgo_world::Cpg& c = w->cpg;
    c = tgo_world::Cpg();
    c.nNodes = d->n_nodes;
    c.begin.assign(d->coupling_begin, d->coupling_begin + d->n_nodes + 1);
    c.target.assign(d->coupling_target, d->coupling_target + d->n_couplings);
    c.pair.assign(d->coupling_pair, d->coupling_pair + d->n_couplings);
    c.impedance.assign(d->impedance, d->impedance + 4 * d->n_nodes);
    c.rConst = d->r_const; c.dMin = d->d_min; c.dMax = d->d_max; c.descCom = d->descending_command;
    c.controlTime = d->control_time; c.cableStep = d->cable_control_step; c.maxSteps = d->max_steps; c.scoreBody = d->score_body;
    c.heightLo = d->height_lo; c.heightHi = d->height_hi;
    c.freqParam = params[0]; c.radiusParam = params[1];
    c.pairs.assign(params + 2, params + 2 + 2 * d->n_pairs);
    c.x.assign(3 * d->n_nodes, 0.0);
    c.nodeValue.assign(d->n_nodes, 0.0);
    c.cableTime.assign(d->n_nodes, 0.0);
    c.cableControl = d->cable_control; c.kP = d->pid_kp; c.kI = d->pid_ki; c.kD = d->pid_kd;
    c.pidPrevError.assign(d->n_nodes, 0.0); c.pidIntError.assign(d->n_nodes, 0.0); c.commandedTension.assign(d->n_nodes, 0.0);
    w->ctrlKind = TGB_CTRL_CPG_IMPEDANCE;
    w->status = 0;
    w->episodeStartCom = w->scoreCom();
    return 0;
}
int tgo_num_contacts(const tgo_world* w) { return (int)w->contacts.size(); }
int tgo_num_pair_contacts(const tgo_world* w) { return w->pairRows; }
void tgo_set_pair_contact_cap(tgo_world* w, int cap) { w->pairCap = cap < 0 ? 0 : cap; }
void tgo_set_warm_start(tgo_world* w, double factor) { w->warmStart = factor; w->manifoldCache.clear(); }
double tgo_get_min_clearance(const tgo_world* w) { return w->minClearance; }
void tgo_reset_min_clearance(tgo_world* w) { w->minClearance = 1e30; }

/* contact cables added to every world tgo_rollout_batch creates (body indices relative to its model); n = 0 clears */
static std::vector<tgb_contact_cable_desc> g_batchContactCables;
void tgo_set_batch_contact_cables(const tgb_contact_cable_desc* cables, int n) {
    g_batchContactCables.assign(cables, cables + (n > 0 ? n : 0));
}

int tgo_rollout_batch(const tgb_world_cfg* cfg, const tgb_ground_desc* ground, const tgb_body_desc* bodies, int n_bodies,
                      const tgb_shape_desc* shapes, int n_shapes, const tgb_cable_desc* cables, int n_cables, int ctrl_kind,
                      const double* ctrl_params, int params_per_env, const double* setup_input, double setup_dt, int n_env,
                      const double* rod_state_in, const double* cable_state_in, double dt, int n_steps, int n_threads,
                      double* rod_state_out, double* cable_obs_out, const tgb_cpg_desc* cpg_desc, double* scores_out, int* status_out) {
    if (!(dt > 0.0)) return TGB_ERR_INVALID_ARGUMENT;
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int t) {
        for (int e = t; e < n_env; e += n_threads) {
            tgo_world* w = tgo_create(cfg, ground);
            tgo_add_model_flat(w, bodies, n_bodies, shapes, n_shapes, cables, n_cables);
            if (!g_batchContactCables.empty()) tgo_add_contact_cables(w, g_batchContactCables.data(), (int)g_batchContactCables.size());
            if (rod_state_in) tgo_set_rod_state(w, rod_state_in + (size_t)e * n_bodies * 13);
            if (cable_state_in) tgo_set_cable_state(w, cable_state_in + (size_t)e * n_cables * 2);
            if (setup_input) tgo_set_control_input(w, setup_input, setup_dt);
            if (ctrl_kind == TGB_CTRL_CPG_IMPEDANCE) tgo_set_cpg_controller(w, cpg_desc, ctrl_params + (size_t)e * params_per_env, params_per_env);
            else if (ctrl_kind != TGB_CTRL_NONE) tgo_set_controller(w, ctrl_kind, ctrl_params + (size_t)e * params_per_env, params_per_env);
            tgo_step(w, dt, n_steps);
            if (rod_state_out) tgo_get_rod_state(w, rod_state_out + (size_t)e * n_bodies * 13);
            if (cable_obs_out) tgo_get_cable_obs(w, cable_obs_out + (size_t)e * n_cables * 4);
            if (scores_out) tgo_get_scores(w, scores_out + (size_t)e * 2);
            if (status_out) status_out[e] = tgo_get_status(w);
            tgo_destroy(w);
        }
    };
    if (n_threads == 1) { work(0); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
    return 0;
}

} // extern "C"
