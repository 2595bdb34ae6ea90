/* Inert stubs (oracle/ref shim) for the sliding-anchor methods that share a translation unit with the pin-anchor code
   (tgBulletSpringCableAnchor.cpp:89-334). Pin anchors pass manifold = NULL, so none of this is ever executed. */
#ifndef REF_SHIM_BT_PERSISTENT_MANIFOLD_H
#define REF_SHIM_BT_PERSISTENT_MANIFOLD_H
#include "LinearMath/btVector3.h"
class btCollisionObject;
class btManifoldPoint {
public:
    btVector3 m_positionWorldOnA, m_positionWorldOnB, m_normalWorldOnB;
    btScalar m_distance1;
    btManifoldPoint() : m_distance1(0) {}
    btScalar getDistance() const { return m_distance1; }
    const btVector3& getPositionWorldOnA() const { return m_positionWorldOnA; }
    const btVector3& getPositionWorldOnB() const { return m_positionWorldOnB; }
};
class btPersistentManifold {
public:
    btPersistentManifold() : m_body0(0), m_body1(0) {}
    const btCollisionObject* getBody0() const { return m_body0; }
    const btCollisionObject* getBody1() const { return m_body1; }
    int getNumContacts() const { return 0; }
    const btManifoldPoint& getContactPoint(int) const { return m_point; }
    btManifoldPoint& getContactPoint(int) { return m_point; }
    const btCollisionObject* m_body0;
    const btCollisionObject* m_body1;
    btManifoldPoint m_point;
};
#endif
