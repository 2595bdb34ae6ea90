/* Shim (oracle/ref, TEST INFRASTRUCTURE) of the Bullet collision-world surface that src/core/tgBulletContactSpringCable.cpp touches: a
   ghost object with its overlapping-pair cache, the broadphase's pair cache with one collision algorithm per pair that hands out
   contact manifolds, a compound shape of cylinders, a dynamics world that owns broadphase and dispatcher.  Nothing here detects
   anything: the harness (ref_harness.cpp) fills the manifolds with the contact points the oracle's geometric restatement finds, so that
   the REFERENCE's own anchor logic runs on the same points as the oracle's restatement of it. */
#ifndef REF_COLLISION_SHIM_H
#define REF_COLLISION_SHIM_H
#include "LinearMath/btTransform.h"
#include "BulletDynamics/Dynamics/btRigidBody.h"
#include "BulletCollision/CollisionShapes/btCollisionShape.h"
#include "BulletCollision/NarrowPhaseCollision/btPersistentManifold.h"
#include "LinearMath/btAlignedObjectArray.h"
#include <cmath>
#include <vector>

inline btScalar btFabs(btScalar x) { return std::fabs(x); }

class btCompoundShape : public btCollisionShape {
public:
    explicit btCompoundShape(bool = true) {}
    void addChildShape(const btTransform&, btCollisionShape* s) { m_children.push_back(s); }
    int getNumChildShapes() const { return (int)m_children.size(); }
    btCollisionShape* getChildShape(int i) { return m_children[i]; }
    void removeChildShapeByIndex(int i) { m_children.erase(m_children.begin() + i); }
private:
    std::vector<btCollisionShape*> m_children;
};
class btCylinderShape : public btCollisionShape {
public:
    explicit btCylinderShape(const btVector3& halfExtents) : m_halfExtents(halfExtents) {}
    btVector3 m_halfExtents;
};
class btBoxShape : public btCollisionShape {
public:
    explicit btBoxShape(const btVector3& halfExtents) : m_halfExtents(halfExtents) {}
    btVector3 m_halfExtents;
};

struct btBroadphaseProxy {
    enum CollisionFilterGroups { DefaultFilter = 1, StaticFilter = 2, CharacterFilter = 32 };
    void* m_clientObject;
    btBroadphaseProxy() : m_clientObject(0) {}
};
typedef btAlignedObjectArray<btPersistentManifold*> btManifoldArray;
class btCollisionAlgorithm {
public:
    void getAllContactManifolds(btManifoldArray& out) { for (size_t i = 0; i < m_manifolds.size(); ++i) out.push_back(m_manifolds[i]); }
    std::vector<btPersistentManifold*> m_manifolds;
};
struct btBroadphasePair {
    btBroadphaseProxy* m_pProxy0;
    btBroadphaseProxy* m_pProxy1;
    btCollisionAlgorithm* m_algorithm;
    btBroadphasePair() : m_pProxy0(0), m_pProxy1(0), m_algorithm(0) {}
};
typedef btAlignedObjectArray<btBroadphasePair> btBroadphasePairArray;
class btDispatcher {};
class btOverlappingPairCache {
public:
    btBroadphasePairArray& getOverlappingPairArray() { return m_pairs; }
    btBroadphasePair* findPair(btBroadphaseProxy* p0, btBroadphaseProxy* p1) {
        for (int i = 0; i < m_pairs.size(); ++i)
            if ((m_pairs[i].m_pProxy0 == p0 && m_pairs[i].m_pProxy1 == p1) || (m_pairs[i].m_pProxy0 == p1 && m_pairs[i].m_pProxy1 == p0)) return &m_pairs[i];
        return 0;
    }
    void cleanProxyFromPairs(btBroadphaseProxy*, btDispatcher*) {}
    btBroadphasePairArray m_pairs;
};
typedef btOverlappingPairCache btHashedOverlappingPairCache;
class btBroadphaseInterface {
public:
    btOverlappingPairCache* getOverlappingPairCache() { return &m_cache; }
    btOverlappingPairCache m_cache;
};
class btGhostObject : public btCollisionObject {
public:
    btGhostObject() : m_shape(0) { m_xf.setIdentity(); m_proxy.m_clientObject = static_cast<btCollisionObject*>(this); }
    btCollisionShape* getCollisionShape() { return m_shape; }
    const btCollisionShape* getCollisionShape() const { return m_shape; }
    void setCollisionShape(btCollisionShape* s) { m_shape = s; }
    void setWorldTransform(const btTransform& t) { m_xf = t; }
    void setCollisionFlags(int) {}
    btBroadphaseProxy* getBroadphaseHandle() { return &m_proxy; }
    btCollisionShape* m_shape;
    btTransform m_xf;
    btBroadphaseProxy m_proxy;
};
class btPairCachingGhostObject : public btGhostObject {
public:
    /* the pairs this ghost overlaps: one per rigid body of the harness world, kept for the cable's lifetime */
    btOverlappingPairCache* getOverlappingPairCache() { return &m_cache; }
    btOverlappingPairCache m_cache;
};
class btDynamicsWorld {
public:
    btBroadphaseInterface* getBroadphase() { return &m_broadphase; }
    btDispatcher* getDispatcher() { return &m_dispatcher; }
    void removeCollisionObject(btCollisionObject*) {}
    btBroadphaseInterface m_broadphase;
    btDispatcher m_dispatcher;
};
typedef btDynamicsWorld btDiscreteDynamicsWorld;
#endif
