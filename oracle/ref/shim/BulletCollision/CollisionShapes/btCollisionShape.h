/* Inert stub (oracle/ref shim): the anchor source includes this header but the pin-anchor path never uses a shape. */
#ifndef REF_SHIM_BT_COLLISION_SHAPE_H
#define REF_SHIM_BT_COLLISION_SHAPE_H
class btCollisionShape {
public:
    virtual ~btCollisionShape() {}
};
#endif
