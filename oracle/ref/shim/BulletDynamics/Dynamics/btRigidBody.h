/*
 * Shim of the btRigidBody surface that the NTRTsim cable / anchor / actuator sources touch.
 * TEST INFRASTRUCTURE (oracle/ref): lets those reference sources be compiled unchanged, in place, without Bullet.
 * Semantics follow Bullet 2.82: applyImpulse(J, r): v += J * invMass; w += invInertiaWorld * (r x J), linear and
 * angular factors 1; a body with invMass == 0 is static and ignores impulses.
 */
#ifndef REF_SHIM_BT_RIGID_BODY_H
#define REF_SHIM_BT_RIGID_BODY_H

#include "LinearMath/btTransform.h"

/* abstractMarker.cpp:31 silences an unused-variable warning through this Bullet global */
static int btInfinityMask = 0x7F800000;

class btCollisionShape;

class btMotionState {
public:
    virtual ~btMotionState() {}
    virtual void getWorldTransform(btTransform& t) const = 0;
};

class btCollisionObject {
public:
    virtual ~btCollisionObject() {}
};

class btRigidBody : public btCollisionObject {
public:
    btRigidBody() : m_invMass(0.0), m_activations(0) {
        m_xf.setIdentity();
        m_invInertiaWorld.setIdentity();
    }
    const btTransform& getWorldTransform() const { return m_xf; }
    btTransform& getWorldTransform() { return m_xf; }
    const btTransform& getCenterOfMassTransform() const { return m_xf; }
    const btVector3& getCenterOfMassPosition() const { return m_xf.getOrigin(); }
    btQuaternion getOrientation() const { return m_xf.getRotation(); }
    btScalar getInvMass() const { return m_invMass; }
    const btVector3& getLinearVelocity() const { return m_linvel; }
    const btVector3& getAngularVelocity() const { return m_angvel; }
    void activate(bool = false) const { ++m_activations; }
    void applyImpulse(const btVector3& impulse, const btVector3& rel_pos) {
        if (m_invMass != btScalar(0.)) {
            m_linvel += impulse * m_invMass;
            m_angvel += m_invInertiaWorld * rel_pos.cross(impulse);
        }
    }
    static const btRigidBody* upcast(const btCollisionObject* o) { return dynamic_cast<const btRigidBody*>(o); }
    static btRigidBody* upcast(btCollisionObject* o) { return dynamic_cast<btRigidBody*>(o); }

    /* harness access */
    btTransform m_xf;
    btVector3 m_linvel, m_angvel;
    btScalar m_invMass;
    btMatrix3x3 m_invInertiaWorld;
    mutable long m_activations;
};

#endif
