/*
 * ref_harness.cpp -- drives the REFERENCE's own cable / anchor / actuator / controller / CPG sources, compiled unchanged
 * and in place from /root/reference/src (see build_ref.py), against the shim headers in oracle/ref/shim.
 * TEST INFRASTRUCTURE: its only purpose is to validate oracle/tg_oracle.cpp and to generate tests/golden/*.json.
 * The world step (Bullet) is not part of this harness; callers move the shim rigid bodies themselves.
 */
#include "core/tgBasicActuator.h"
#include "core/tgKinematicActuator.h"
#include "core/tgBulletSpringCable.h"
#include "core/tgBulletSpringCableAnchor.h"
#include "core/tgSpringCableActuator.h"
#include "controllers/tgTensionController.h"
#include "controllers/tgImpedanceController.h"
#include "controllers/tgPIDController.h"
#include "util/CPGEquations.h"
#include "BulletDynamics/Dynamics/btRigidBody.h"
#include "core/tgBulletContactSpringCable.h"
#include "core/tgBulletUtil.h"
#include "core/tgWorld.h"
#include "ref_collision_shim.h"

#include <stdexcept>
#include <vector>

/* the one Bullet world of the harness: tgBulletContactSpringCable asks tgBulletUtil for it (tgBulletContactSpringCable.cpp:76,206,682) */
static btDynamicsWorld g_dynamicsWorld;
btDynamicsWorld& tgBulletUtil::worldToDynamicsWorld(const tgWorld&) { return g_dynamicsWorld; }

/* a contact cable of the reference with the collision-world objects its step reads: the ghost, one broadphase pair per rigid body,
   each with an algorithm that holds one manifold (body0 = the ghost, body1 = the rigid body) */
struct ref_contact_cable {
    tgBulletContactSpringCable* cable;
    tgBasicActuator* actuator;
    btPairCachingGhostObject* ghost;
    std::vector<btBroadphaseProxy*> proxies;
    std::vector<btCollisionAlgorithm*> algorithms;
    std::vector<btPersistentManifold*> manifolds;
};

struct ref_world {
    std::vector<ref_contact_cable> contactCables;
    std::vector<btRigidBody*> bodies;
    std::vector<btVector3> invInertiaLocal;
    std::vector<tgSpringCableActuator*> actuators;
    std::vector<tgBulletSpringCable*> cables;
    std::vector<tgTensionController*> tensionControllers;
    std::vector<tgPIDController*> pids;
    std::vector<int> kinds;
    ~ref_world() {
        for (size_t i = 0; i < tensionControllers.size(); ++i) delete tensionControllers[i];
        for (size_t i = 0; i < pids.size(); ++i) delete pids[i];
        for (size_t i = 0; i < actuators.size(); ++i) delete actuators[i]; /* deletes its cable and the anchors */
        for (size_t i = 0; i < contactCables.size(); ++i) {
            /* the world's pair cache outlives this harness world: take this cable's pairs out before its objects go */
            btBroadphasePairArray& all = g_dynamicsWorld.getBroadphase()->getOverlappingPairCache()->getOverlappingPairArray();
            btBroadphasePairArray kept;
            for (int k = 0; k < all.size(); ++k)
                if (all[k].m_pProxy0 != contactCables[i].ghost->getBroadphaseHandle()) kept.push_back(all[k]);
            all = kept;
            delete contactCables[i].actuator; /* deletes the cable, which deletes its ghost object and shapes */
            for (size_t k = 0; k < contactCables[i].proxies.size(); ++k) {
                delete contactCables[i].proxies[k]; delete contactCables[i].algorithms[k]; delete contactCables[i].manifolds[k];
            }
        }
        for (size_t i = 0; i < bodies.size(); ++i) delete bodies[i];
    }
};

static void setBody(ref_world* w, int b, const double* s) {
    btRigidBody* rb = w->bodies[b];
    rb->m_xf.setOrigin(btVector3(s[0], s[1], s[2]));
    rb->m_xf.setRotation(btQuaternion(s[3], s[4], s[5], s[6]));
    rb->m_linvel = btVector3(s[7], s[8], s[9]);
    rb->m_angvel = btVector3(s[10], s[11], s[12]);
    /* btRigidBody::updateInertiaTensor */
    rb->m_invInertiaWorld = rb->m_xf.getBasis().scaled(w->invInertiaLocal[b]) * rb->m_xf.getBasis().transpose();
}

extern "C" {

ref_world* ref_create() { return new ref_world(); }
void ref_destroy(ref_world* w) { delete w; }

int ref_add_body(ref_world* w, const double* state13, double mass, const double* inertia) {
    btRigidBody* rb = new btRigidBody();
    rb->m_invMass = mass != 0.0 ? 1.0 / mass : 0.0;
    w->bodies.push_back(rb);
    w->invInertiaLocal.push_back(mass != 0.0 ? btVector3(inertia[0] != 0.0 ? 1.0 / inertia[0] : 0.0, inertia[1] != 0.0 ? 1.0 / inertia[1] : 0.0,
                                                          inertia[2] != 0.0 ? 1.0 / inertia[2] : 0.0)
                                             : btVector3(0, 0, 0));
    setBody(w, (int)w->bodies.size() - 1, state13);
    return (int)w->bodies.size() - 1;
}

void ref_set_body(ref_world* w, int b, const double* state13) { setBody(w, b, state13); }

void ref_get_body(ref_world* w, int b, double* s) {
    btRigidBody* rb = w->bodies[b];
    const btVector3& o = rb->m_xf.getOrigin();
    btQuaternion q = rb->m_xf.getRotation();
    s[0] = o.x(); s[1] = o.y(); s[2] = o.z();
    s[3] = q.x(); s[4] = q.y(); s[5] = q.z(); s[6] = q.w();
    s[7] = rb->m_linvel.x(); s[8] = rb->m_linvel.y(); s[9] = rb->m_linvel.z();
    s[10] = rb->m_angvel.x(); s[11] = rb->m_angvel.y(); s[12] = rb->m_angvel.z();
}

/* kind 1 = tgBasicActuator, 2 = tgKinematicActuator. cfg = {k, c, pretension, maxTens, tVel, minAL, minRL, radius, friction, inertia, backdrivable}.
   Returns the cable index or -1 - and the exception text through ref_last_error - when the reference constructor throws. */
static std::string g_err;
const char* ref_last_error() { return g_err.c_str(); }

int ref_add_cable(ref_world* w, int a, int b, const double* worldA, const double* worldB, int kind, const double* cfg) {
    try {
        std::vector<tgBulletSpringCableAnchor*> anchors;
        anchors.push_back(new tgBulletSpringCableAnchor(w->bodies[a], btVector3(worldA[0], worldA[1], worldA[2])));
        anchors.push_back(new tgBulletSpringCableAnchor(w->bodies[b], btVector3(worldB[0], worldB[1], worldB[2])));
        tgBulletSpringCable* cable = new tgBulletSpringCable(anchors, cfg[0], cfg[1], cfg[2]);
        tgSpringCableActuator* act = 0;
        if (kind == 2) {
            tgKinematicActuator::Config kc(cfg[0], cfg[1], cfg[2], cfg[7], cfg[8], cfg[9], cfg[10] != 0.0, false, cfg[3], cfg[4], cfg[5], cfg[6]);
            act = new tgKinematicActuator(cable, tgTags("muscle"), kc);
        } else {
            tgSpringCableActuator::Config bc(cfg[0], cfg[1], cfg[2], false, cfg[3], cfg[4], cfg[5], cfg[6]);
            act = new tgBasicActuator(cable, tgTags("muscle"), bc);
        }
        w->cables.push_back(cable);
        w->actuators.push_back(act);
        w->kinds.push_back(kind);
        w->tensionControllers.push_back(kind == 1 ? new tgTensionController(static_cast<tgBasicActuator*>(act), 0.0) : 0);
        w->pids.push_back(0);
        return (int)w->actuators.size() - 1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

/* out = {actuator rest length, cable rest length, actual length, tension, cable velocity, actuator getVelocity, start length, damping} */
void ref_get_cable(ref_world* w, int c, double* out) {
    out[0] = w->actuators[c]->getRestLength();
    out[1] = w->cables[c]->getRestLength();
    out[2] = w->actuators[c]->getCurrentLength();
    out[3] = w->actuators[c]->getTension();
    out[4] = w->cables[c]->getVelocity();
    out[5] = w->actuators[c]->getVelocity();
    out[6] = w->actuators[c]->getStartLength();
    out[7] = w->cables[c]->getDamping();
}

/* tgBasicActuator::setControlInput(input, dt); 1 = std::invalid_argument */
int ref_set_control_input_dt(ref_world* w, int c, double input, double dt) {
    try {
        w->actuators[c]->setControlInput(input, dt);
        return 0;
    } catch (const std::invalid_argument& e) { g_err = e.what(); return 1; }
}
/* tgControllable::setControlInput(input): preferred length (basic) or desired torque (kinematic) */
int ref_set_control_input(ref_world* w, int c, double input) {
    try {
        w->actuators[c]->setControlInput(input);
        return 0;
    } catch (const std::invalid_argument& e) { g_err = e.what(); return 1; }
}
void ref_move_motors(ref_world* w, int c, double dt) {
    if (w->kinds[c] == 1) static_cast<tgBasicActuator*>(w->actuators[c])->moveMotors(dt);
}
/* T6TensionController::onStep body for one actuator: tgTensionController::control(dt, setPoint) */
int ref_tension_control(ref_world* w, int c, double dt, double setPoint) {
    try {
        w->tensionControllers[c]->control(dt, setPoint);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 5; }
}
/* tgImpedanceController(T0, Kp, Kv).control(tgBasicActuator&, dt, newPosition, offsetVel); returns the set tension */
double ref_impedance_control(ref_world* w, int c, double dt, double T0, double Kp, double Kv, double newPosition, double offsetVel) {
    tgImpedanceController ic(T0, Kp, Kv);
    return ic.control(*static_cast<tgBasicActuator*>(w->actuators[c]), dt, newPosition, offsetVel);
}
/* tgCPGCableControl::onSetup: m_PID = new tgPIDController(&subject, Config(p, i, d, tensControl)) (tgCPGCableControl.cpp:52-61) */
int ref_pid_create(ref_world* w, int c, double p, double i, double d, int tensControl) {
    try {
        delete w->pids[c];
        w->pids[c] = new tgPIDController(w->actuators[c], tgPIDController::Config(p, i, d, tensControl != 0));
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
/* tgImpedanceController(T0, Kp, Kv).control(tgBasicController& = the cable's tgPIDController, dt, newPosition, offsetVel)
   (tgCPGCableControl.cpp:73); returns the set tension */
double ref_impedance_control_pid(ref_world* w, int c, double dt, double T0, double Kp, double Kv, double newPosition, double offsetVel) {
    tgImpedanceController ic(T0, Kp, Kv);
    return ic.control(*w->pids[c], dt, newPosition, offsetVel);
}
/* m_PID->control(dt, commandedTension, currentTension) with currentTension = subject.getTension() (tgCPGCableControl.cpp:83-87) */
void ref_pid_control(ref_world* w, int c, double dt, double setPoint) {
    const double currentTension = w->actuators[c]->getTension();
    w->pids[c]->control(dt, setPoint, currentTension);
}
/* tgModel::step over the connectors: every actuator's step(dt) in creation order; 1 = std::invalid_argument */
int ref_step_actuators(ref_world* w, double dt) {
    try {
        for (size_t i = 0; i < w->actuators.size(); ++i) w->actuators[i]->step(dt);
        return 0;
    } catch (const std::invalid_argument& e) { g_err = e.what(); return 1; }
}

/* ---- contact cables: src/core/tgBulletContactSpringCable.cpp compiled in place ------------------------------ */
/* cfg = {k, c, pretension, thickness, resolution}; the ends as world points on bodies a and b */
int ref_add_contact_cable(ref_world* w, int a, int b, const double* worldA, const double* worldB, const double* cfg) {
    try {
        static char worldStorage[4096]; /* a tgWorld is only ever handed to tgBulletUtil::worldToDynamicsWorld above */
        tgWorld& world = *reinterpret_cast<tgWorld*>(worldStorage);
        ref_contact_cable cc;
        cc.ghost = new btPairCachingGhostObject();
        cc.ghost->setCollisionShape(new btCompoundShape());
        for (size_t i = 0; i < w->bodies.size(); ++i) {
            btBroadphaseProxy* proxy = new btBroadphaseProxy();
            proxy->m_clientObject = static_cast<btCollisionObject*>(w->bodies[i]);
            btPersistentManifold* m = new btPersistentManifold();
            m->m_body0 = cc.ghost; m->m_body1 = w->bodies[i];
            btCollisionAlgorithm* alg = new btCollisionAlgorithm();
            alg->m_manifolds.push_back(m);
            btBroadphasePair pair;
            pair.m_pProxy0 = cc.ghost->getBroadphaseHandle(); pair.m_pProxy1 = proxy; pair.m_algorithm = alg;
            cc.ghost->getOverlappingPairCache()->getOverlappingPairArray().push_back(pair);
            g_dynamicsWorld.getBroadphase()->getOverlappingPairCache()->getOverlappingPairArray().push_back(pair);
            cc.proxies.push_back(proxy); cc.algorithms.push_back(alg); cc.manifolds.push_back(m);
        }
        std::vector<tgBulletSpringCableAnchor*> anchors;
        anchors.push_back(new tgBulletSpringCableAnchor(w->bodies[a], btVector3(worldA[0], worldA[1], worldA[2])));
        anchors.push_back(new tgBulletSpringCableAnchor(w->bodies[b], btVector3(worldB[0], worldB[1], worldB[2])));
        cc.cable = new tgBulletContactSpringCable(cc.ghost, world, anchors, cfg[0], cfg[1], cfg[2], cfg[3], cfg[4]);
        tgSpringCableActuator::Config bc(cfg[0], cfg[1], cfg[2], false, 600000000, cfg[5] > 0.0 ? cfg[5] : 100.0);
        cc.actuator = new tgBasicActuator(cc.cable, tgTags("muscle"), bc); /* tgBasicContactCableInfo::createModel */
        w->contactCables.push_back(cc);
        return (int)w->contactCables.size() - 1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
/* the cable's anchors in order: world positions [n][3]; returns n */
int ref_cc_chain(ref_world* w, int c, double* out, int cap) {
    const std::vector<const tgSpringCableAnchor*> anchors = w->contactCables[c].cable->getAnchors();
    for (int i = 0; i < (int)anchors.size() && i < cap; ++i) {
        const btVector3 p = anchors[i]->getWorldPosition();
        out[3 * i] = p.x(); out[3 * i + 1] = p.y(); out[3 * i + 2] = p.z();
    }
    return (int)anchors.size();
}
/* this step's contact points of the cable, as the oracle's geometric manifold finds them: pts [n][8] = body, position on the body
   (3), normal from the cable into the body (3), distance.  Bullet's convention for a manifold with the ghost as body 0:
   m_positionWorldOnB on the rigid body, m_normalWorldOnB from the rigid body towards the ghost. */
void ref_cc_set_points(ref_world* w, int c, int n, const double* pts) {
    ref_contact_cable& cc = w->contactCables[c];
    for (size_t i = 0; i < cc.manifolds.size(); ++i) cc.manifolds[i]->m_points.clear();
    for (int i = 0; i < n; ++i) {
        const double* p = pts + 8 * i;
        btManifoldPoint pt;
        pt.m_positionWorldOnB = btVector3(p[1], p[2], p[3]);
        pt.m_normalWorldOnB = btVector3(-p[4], -p[5], -p[6]);
        pt.m_positionWorldOnA = pt.m_positionWorldOnB + pt.m_normalWorldOnB * p[7];
        pt.m_distance1 = p[7];
        cc.manifolds[(int)p[0]]->m_points.push_back(pt);
    }
}
/* tgBasicActuator::step(dt) of the contact cable's actuator = tgBulletContactSpringCable::step; 5 = std::runtime_error */
int ref_cc_step(ref_world* w, int c, double dt) {
    try {
        w->contactCables[c].actuator->step(dt);
        return 0;
    } catch (const std::invalid_argument& e) { g_err = e.what(); return 1;
    } catch (const std::exception& e) { g_err = e.what(); return 5; }
}
/* tgBasicActuator::setControlInput(input, dt) on the contact cable's actuator (moveMotors on the length over all anchors);
   1 = std::invalid_argument */
int ref_cc_set_control_input_dt(ref_world* w, int c, double input, double dt) {
    try {
        w->contactCables[c].actuator->setControlInput(input, dt);
        return 0;
    } catch (const std::invalid_argument& e) { g_err = e.what(); return 1; }
}
/* out = {rest length, actual length, tension, number of anchors} */
void ref_cc_get(ref_world* w, int c, double* out) {
    tgBulletContactSpringCable* cable = w->contactCables[c].cable;
    out[0] = w->contactCables[c].actuator->getRestLength(); out[1] = cable->getActualLength(); out[2] = w->contactCables[c].actuator->getTension();
    out[3] = (double)cable->getAnchors().size();
}
/* the sliding anchors: [k][7] = index of the attached body, world position, contact normal; returns their number */
int ref_cc_anchors(ref_world* w, int c, double* out, int cap) {
    const std::vector<const tgSpringCableAnchor*> anchors = w->contactCables[c].cable->getAnchors();
    int k = 0;
    for (size_t i = 1; i + 1 < anchors.size() && k < cap; ++i, ++k) {
        const tgBulletSpringCableAnchor* a = static_cast<const tgBulletSpringCableAnchor*>(anchors[i]);
        int body = -1;
        for (size_t b = 0; b < w->bodies.size(); ++b) if (w->bodies[b] == a->attachedBody) body = (int)b;
        const btVector3 p = a->getWorldPosition(), n = a->getContactNormal();
        double* o = out + 7 * k;
        o[0] = body; o[1] = p.x(); o[2] = p.y(); o[3] = p.z(); o[4] = n.x(); o[5] = n.y(); o[6] = n.z();
    }
    return k;
}

/* ---- CPG ------------------------------------------------------------------------------------------------- */
CPGEquations* ref_cpg_create(int maxSteps) { return new CPGEquations(maxSteps); }
void ref_cpg_destroy(CPGEquations* c) { delete c; }
int ref_cpg_add_node(CPGEquations* c, const double* params7) {
    std::vector<double> p(params7, params7 + 7);
    return c->addNode(p);
}
void ref_cpg_define_connections(CPGEquations* c, int node, int n, const int* conn, const double* weights, const double* phases) {
    c->defineConnections(node, std::vector<int>(conn, conn + n), std::vector<double>(weights, weights + n), std::vector<double>(phases, phases + n));
}
/* 0 ok, 5 = std::runtime_error ("Inefficient CPG Parameters") */
int ref_cpg_update(CPGEquations* c, const double* descCom, int n, double dt) {
    std::vector<double> d(descCom, descCom + n);
    try {
        c->update(d, dt);
        return 0;
    } catch (const std::runtime_error& e) { g_err = e.what(); return 5; }
}
double ref_cpg_value(CPGEquations* c, int i) { return (*c)[i]; }
void ref_cpg_state(CPGEquations* c, double* x) {
    std::vector<double>& v = c->getXVars();
    for (size_t i = 0; i < v.size(); ++i) x[i] = v[i];
}

} /* extern "C" */
