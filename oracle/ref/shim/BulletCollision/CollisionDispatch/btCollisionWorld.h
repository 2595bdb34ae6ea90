/* Shim (oracle/ref). */
#ifndef REF_SHIM_BT_COLLISION_WORLD_H
#define REF_SHIM_BT_COLLISION_WORLD_H
#include "BulletCollision/NarrowPhaseCollision/btPersistentManifold.h"
class btCollisionWorld {};
#endif
