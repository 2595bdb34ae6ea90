/* Shim (oracle/ref) of the part of btPersistentManifold the sliding-anchor / contact-cable sources read
   (tgBulletSpringCableAnchor.cpp:89-334, tgBulletContactSpringCable.cpp:196-365): the two bodies and a list of contact points.  The
   harness fills the points (ref_harness.cpp ref_cc_set_points); pin anchors pass manifold = NULL and never get here. */
#ifndef REF_SHIM_BT_PERSISTENT_MANIFOLD_H
#define REF_SHIM_BT_PERSISTENT_MANIFOLD_H
#include "LinearMath/btVector3.h"
#include <vector>
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
    int getNumContacts() const { return (int)m_points.size(); }
    const btManifoldPoint& getContactPoint(int i) const { return m_points[i]; }
    btManifoldPoint& getContactPoint(int i) { return m_points[i]; }
    const btCollisionObject* m_body0;
    const btCollisionObject* m_body1;
    std::vector<btManifoldPoint> m_points;
};
#endif
