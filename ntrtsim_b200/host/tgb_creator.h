/*
 * tgb_creator.h -- the "tgcreator" half of the host API: nodes, pairs, tags, build spec and the
 * structure-to-bodies flattener.  Names and call sequences follow src/tgcreator of the reference
 * (tgStructure.h, tgBuildSpec.h, tgRodInfo.h, tgBasicActuatorInfo.h, tgStructureInfo.h, tgUtil.h);
 * the implementation is new and emits tgb_body_desc / tgb_shape_desc / tgb_cable_desc records into the
 * tgWorldB200Impl instead of creating Bullet objects.
 */
#ifndef TGB_CREATOR_H
#define TGB_CREATOR_H

#include "tgb_host.h"

#include <cmath>

/* geometry helpers with the reference's arithmetic (src/tgcreator/tgUtil.h:92-311) */
struct tgUtil {
    static btVector3 upVector() { return btVector3(0.0, 1.0, 0.0); }
    static btVector3 center(const btVector3& a, const btVector3& b) { return (a + b) / 2.0; }
    static bool almostEqual(const btVector3& a, const btVector3& b, int precision = 5);
    static btQuaternion getQuaternionBetween(btVector3 a, btVector3 b, const btVector3& fallbackAxis = btVector3(0, 0, 0));
    static btTransform getTransform(const btVector3& start, const btVector3& end);
    static void addRotation(btVector3& v, const btVector3& fixedPoint, const btVector3& axis, double angle);
    static void addRotation(btVector3& v, const btVector3& fixedPoint, const btQuaternion& rotation);
    static void addRotation(btVector3& v, const btVector3& fixedPoint, const btVector3& fromOrientation, const btVector3& toOrientation) {
        addRotation(v, fixedPoint, getQuaternionBetween(fromOrientation, toOrientation));
    }
    /* tgUtil.h:324-365: unit conversion, rounding to `precision` decimals, the C library's generator seeded (from the clock / a number) */
    static double rad2deg(double radians) { return radians * 180.0 / M_PI; }
    static double deg2rad(double degrees) { return degrees * (M_PI / 180.0); }
    static double round(double d, int precision = 5) {
        const int m = static_cast<int>(pow(10.0, precision));
        return floor(d * m + 0.5) / m;
    }
    static void seedRandom();
    static void seedRandom(int seed);
};

/* src/core/tgString.h: a name with a number appended ("segment", 3 -> "segment3") */
inline std::string tgString(std::string s, int i) {
    std::ostringstream ss;
    ss << s << i;
    return ss.str();
}

/* A tagged point (tgNode.h). */
class tgNode : public btVector3, public tgTaggable {
public:
    tgNode(const btVector3& v = btVector3(), const std::string& tags = "") : btVector3(v), tgTaggable(tgTags(tags)) {}
    tgNode(double x, double y, double z, const std::string& tags = "") : btVector3(x, y, z), tgTaggable(tgTags(tags)) {}
    /* tgNode.h:62-92: turn the node about a fixed point */
    void addRotation(const btVector3& fixedPoint, const btVector3& axis, double angle) { tgUtil::addRotation(*this, fixedPoint, axis, angle); }
    void addRotation(const btVector3& fixedPoint, const btQuaternion& rotation) { addRotation(fixedPoint, rotation.getAxis(), rotation.getAngle()); }
    void addRotation(const btVector3& fixedPoint, const btVector3& fromOrientation, const btVector3& toOrientation) {
        addRotation(fixedPoint, tgUtil::getQuaternionBetween(fromOrientation, toOrientation));
    }
    bool operator==(const tgNode& other) const { return x() == other.x() && y() == other.y() && z() == other.z(); }
};

/* A tagged pair of points (tgPair.h). */
class tgPair : public tgTaggable {
public:
    tgPair() {}
    tgPair(const btVector3& from, const btVector3& to, const std::string& tags = "") : tgTaggable(tgTags(tags)), m_from(from), m_to(to) {}
    btVector3& getFrom() { return m_from; }
    const btVector3& getFrom() const { return m_from; }
    btVector3& getTo() { return m_to; }
    const btVector3& getTo() const { return m_to; }
    void move(const btVector3& offset) { m_from += offset; m_to += offset; }
    void addRotation(const btVector3& fixedPoint, const btQuaternion& rotation);
    void scale(const btVector3& ref, double f);
private:
    btVector3 m_from, m_to;
};

class tgNodes {
public:
    void addNode(double x, double y, double z, const std::string& tags = "") { m_nodes.push_back(tgNode(x, y, z, tags)); }
    void addNode(const tgNode& n) { m_nodes.push_back(n); }
    std::size_t size() const { return m_nodes.size(); }
    tgNode& operator[](std::size_t i);
    const tgNode& operator[](std::size_t i) const;
    void move(const btVector3& offset);
    void addRotation(const btVector3& fixedPoint, const btQuaternion& rotation);
    void scale(const btVector3& ref, double f);
private:
    std::vector<tgNode> m_nodes;
};

class tgPairs {
public:
    void addPair(const tgPair& p) { m_pairs.push_back(p); }
    /* erases every pair with the same end points, tags not compared (tgPair.h:110-113, tgTaggables.h:191-193) */
    void removePair(const tgPair& p);
    std::size_t size() const { return m_pairs.size(); }
    tgPair& operator[](std::size_t i) { return m_pairs.at(i); }
    const tgPair& operator[](std::size_t i) const { return m_pairs.at(i); }
    void move(const btVector3& offset);
    void addRotation(const btVector3& fixedPoint, const btQuaternion& rotation);
    void scale(const btVector3& ref, double f);
private:
    std::vector<tgPair> m_pairs;
};

/* Nodes + pairs + child structures (tgStructure.h). */
class tgStructure : public tgTaggable {
public:
    tgStructure();
    explicit tgStructure(const tgTags& tags);
    explicit tgStructure(const std::string& tags);
    tgStructure(const tgStructure& other);
    tgStructure& operator=(const tgStructure& other);
    virtual ~tgStructure();

    void addNode(double x, double y, double z, std::string tags = "");
    void addNode(tgNode& newNode);
    /* index out of range -> std::out_of_range */
    void addPair(int fromNodeIdx, int toNodeIdx, std::string tags = "");
    void addPair(const btVector3& from, const btVector3& to, std::string tags = "");
    void move(const btVector3& offset);
    void addRotation(const btVector3& fixedPoint, const btVector3& axis, double angle);
    void addRotation(const btVector3& fixedPoint, const btVector3& fromOrientation, const btVector3& toOrientation);
    void addRotation(const btVector3& fixedPoint, const btQuaternion& rotation);
    void scale(double scaleFactor);
    void scale(const btVector3& referencePoint, double scaleFactor);
    void addChild(tgStructure* child);        /* takes ownership */
    void addChild(const tgStructure& child);  /* copies */
    btVector3 getCentroid() const;
    const tgNodes& getNodes() const { return m_nodes; }
    /* breadth-first over this structure and its children; not found -> std::invalid_argument (tgStructure.cpp:206-267) */
    tgNode& findNode(const std::string& tags);
    tgStructure& findChild(const std::string& tags);
    /* either direction; not found -> std::invalid_argument (tgStructure.cpp:226-247) */
    tgPair& findPair(const btVector3& from, const btVector3& to);
    /* here and in every child (tgStructure.cpp:101-106) */
    void removePair(const tgPair& pair);
    const tgPairs& getPairs() const { return m_pairs; }
    const std::vector<tgStructure*>& getChildren() const { return m_children; }
private:
    tgNodes m_nodes;
    tgPairs m_pairs;
    std::vector<tgStructure*> m_children;
};

/* ---- rigid infos ------------------------------------------------------------------------------ */
class tgRigidInfo : public tgTaggable {
public:
    tgRigidInfo() : m_group(0), m_body(-1) {}
    explicit tgRigidInfo(const tgTags& tags) : tgTaggable(tags), m_group(0), m_body(-1) {}
    virtual ~tgRigidInfo() {}
    /* factory side: NULL when the search does not match the pair's / node's tags (tgRigidInfo.h:105-117: a factory answers for
       pairs or for nodes, the other default is NULL) */
    virtual tgRigidInfo* createRigidInfo(const tgPair& pair, const tgTagSearch& search) const { (void)pair; (void)search; return 0; }
    virtual tgRigidInfo* createRigidInfo(const tgNode& node, const tgTagSearch& search) const { (void)node; (void)search; return 0; }
    /* friction, rolling friction, restitution of this rigid's Config */
    virtual void getMaterial(double& friction, double& rollFriction, double& restitution) const = 0;
    virtual double getMass() const = 0;
    virtual btVector3 getCenterOfMass() const = 0;
    virtual btTransform getTransform() const = 0;
    virtual bool containsNode(const btVector3& node) const = 0;
    virtual std::vector<btVector3> getContainedNodes() const = 0;
    virtual btVector3 getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const = 0;
    virtual void appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const = 0;
    virtual tgb_body_desc describeBody(std::vector<tgb_shape_desc>& shapes) const = 0;
    virtual tgModel* createModel(tgWorldB200Impl& world) = 0;
    bool sharesNodesWith(const tgRigidInfo& other) const;
    void setRigidInfoGroup(tgRigidInfo* g) { m_group = g; }
    tgRigidInfo* getRigidInfoGroup() const { return m_group; }
    /* body index in the flattened world, shared by all members of a compound */
    int getBodyIndex() const { return m_group && m_group != this ? m_group->getBodyIndex() : m_body; }
    void setBodyIndex(int b) { m_body = b; }
    void initRigidBody(tgWorldB200Impl& world);
private:
    tgRigidInfo* m_group;
    int m_body;
};

/* A rod between the two points of a pair: flat-ended cylinder (tgRodInfo.cpp:72-177). */
class tgRodInfo : public tgRigidInfo {
public:
    explicit tgRodInfo(const tgRod::Config& config);
    tgRodInfo(const tgRod::Config& config, const tgTags& tags, const tgPair& pair);
    tgRigidInfo* createRigidInfo(const tgPair& pair, const tgTagSearch& search) const;
    void getMaterial(double& friction, double& rollFriction, double& restitution) const {
        friction = m_config.friction; rollFriction = m_config.rollFriction; restitution = m_config.restitution;
    }
    const btVector3& getFrom() const { return m_pair.getFrom(); }
    const btVector3& getTo() const { return m_pair.getTo(); }
    double getLength() const { return getFrom().distance(getTo()); }
    double getMass() const;
    btVector3 getCenterOfMass() const { return tgUtil::center(getFrom(), getTo()); }
    btTransform getTransform() const { return tgUtil::getTransform(getFrom(), getTo()); }
    bool containsNode(const btVector3& node) const { return getFrom() == node || getTo() == node; }
    std::vector<btVector3> getContainedNodes() const;
    btVector3 getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const;
    void appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const;
    tgb_body_desc describeBody(std::vector<tgb_shape_desc>& shapes) const;
    tgModel* createModel(tgWorldB200Impl& world);
    const tgRod::Config& getConfig() const { return m_config; }
private:
    tgPair m_pair;
    tgRod::Config m_config;
};

/* A box between the two points of a pair: btBoxShape(width, length / 2, height) in the pair's frame (tgBoxInfo.cpp:72-200). */
class tgBoxInfo : public tgRigidInfo {
public:
    explicit tgBoxInfo(const tgBox::Config& config);
    tgBoxInfo(const tgBox::Config& config, const tgTags& tags, const tgPair& pair);
    tgRigidInfo* createRigidInfo(const tgPair& pair, const tgTagSearch& search) const;
    void getMaterial(double& friction, double& rollFriction, double& restitution) const {
        friction = m_config.friction; rollFriction = m_config.rollFriction; restitution = m_config.restitution;
    }
    const btVector3& getFrom() const { return m_pair.getFrom(); }
    const btVector3& getTo() const { return m_pair.getTo(); }
    double getLength() const { return getFrom().distance(getTo()); }
    double getMass() const;
    btVector3 getCenterOfMass() const { return tgUtil::center(getFrom(), getTo()); }
    btTransform getTransform() const { return tgUtil::getTransform(getFrom(), getTo()); }
    bool containsNode(const btVector3& node) const { return getFrom() == node || getTo() == node; }
    std::vector<btVector3> getContainedNodes() const;
    btVector3 getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const;
    void appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const;
    tgb_body_desc describeBody(std::vector<tgb_shape_desc>& shapes) const;
    tgModel* createModel(tgWorldB200Impl& world);
    const tgBox::Config& getConfig() const { return m_config; }
private:
    tgb_shape_desc shape() const;
    tgPair m_pair;
    tgBox::Config m_config;
};

/* A sphere at a tagged node (tgSphereInfo.cpp:45-200). */
class tgSphereInfo : public tgRigidInfo {
public:
    explicit tgSphereInfo(const tgSphere::Config& config);
    tgSphereInfo(const tgSphere::Config& config, const tgTags& tags, const tgNode& node);
    tgRigidInfo* createRigidInfo(const tgNode& node, const tgTagSearch& search) const;
    void getMaterial(double& friction, double& rollFriction, double& restitution) const {
        friction = m_config.friction; rollFriction = m_config.rollFriction; restitution = m_config.restitution;
    }
    const btVector3& getNode() const { return m_node; }
    double getMass() const;
    btVector3 getCenterOfMass() const { return m_node; }
    btTransform getTransform() const;
    bool containsNode(const btVector3& node) const { return m_node == node; }
    std::vector<btVector3> getContainedNodes() const { return std::vector<btVector3>(1, m_node); }
    btVector3 getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const;
    void appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const;
    tgb_body_desc describeBody(std::vector<tgb_shape_desc>& shapes) const;
    tgModel* createModel(tgWorldB200Impl& world);
    const tgSphere::Config& getConfig() const { return m_config; }
private:
    tgNode m_node;
    tgSphere::Config m_config;
};

/* Several rigids that share nodes, welded into one body (tgCompoundRigidInfo.cpp:47-112). */
class tgCompoundRigidInfo : public tgRigidInfo {
public:
    tgCompoundRigidInfo() {}
    void addRigid(tgRigidInfo& r) { m_rigids.push_back(&r); }
    void getMaterial(double& friction, double& rollFriction, double& restitution) const;
    double getMass() const;
    btVector3 getCenterOfMass() const;
    btTransform getTransform() const;
    bool containsNode(const btVector3& node) const;
    std::vector<btVector3> getContainedNodes() const;
    btVector3 getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const;
    void appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const;
    tgb_body_desc describeBody(std::vector<tgb_shape_desc>& shapes) const;
    tgModel* createModel(tgWorldB200Impl& world);
private:
    std::vector<tgRigidInfo*> m_rigids;
};

/* ---- connector infos -------------------------------------------------------------------------- */
class tgConnectorInfo : public tgTaggable {
public:
    tgConnectorInfo() : m_fromRigid(0), m_toRigid(0) {}
    tgConnectorInfo(const tgTags& tags, const tgPair& pair) : tgTaggable(tags), m_pair(pair), m_fromRigid(0), m_toRigid(0) {}
    virtual ~tgConnectorInfo() {}
    virtual tgConnectorInfo* createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const = 0;
    virtual void initConnector(tgWorldB200Impl& world) = 0;
    virtual tgModel* createModel(tgWorldB200Impl& world) = 0;
    const btVector3& getFrom() const { return m_pair.getFrom(); }
    const btVector3& getTo() const { return m_pair.getTo(); }
    void chooseRigids(const std::vector<tgRigidInfo*>& rigids);
    tgRigidInfo* getFromRigidInfo() const { return m_fromRigid; }
    tgRigidInfo* getToRigidInfo() const { return m_toRigid; }
protected:
    tgb_cable_desc describeCable(const tgSpringCableActuator::Config& config, tgWorldB200Impl& world, double& startLength) const;
    tgPair m_pair;
    tgRigidInfo* m_fromRigid;
    tgRigidInfo* m_toRigid;
};

class tgBasicActuatorInfo : public tgConnectorInfo {
public:
    explicit tgBasicActuatorInfo(const tgBasicActuator::Config& config);
    tgBasicActuatorInfo(const tgBasicActuator::Config& config, const tgTags& tags, const tgPair& pair);
    tgConnectorInfo* createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const;
    void initConnector(tgWorldB200Impl& world);
    tgModel* createModel(tgWorldB200Impl& world);
private:
    tgBasicActuator::Config m_config;
    int m_cable;
    double m_startLength;
};

/* tgBasicContactCableInfo (src/tgcreator/tgBasicContactCableInfo.cpp:62-131): a tgBasicActuator on a tgBulletContactSpringCable -- a
   cable that collides with the rigid bodies and bends around them over sliding anchors (include/ntrtsim_b200.h "Contact cables").
   Both ends always move to the rigids' connection points (:90-91), thickness 0.001 and resolution 0.1 are the cable's defaults
   (src/core/tgBulletContactSpringCable.h:89-90). */
class tgBasicContactCableInfo : public tgConnectorInfo {
public:
    explicit tgBasicContactCableInfo(const tgBasicActuator::Config& config);
    tgBasicContactCableInfo(const tgBasicActuator::Config& config, const tgTags& tags, const tgPair& pair);
    tgConnectorInfo* createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const;
    void initConnector(tgWorldB200Impl& world);
    tgModel* createModel(tgWorldB200Impl& world);
private:
    tgBasicActuator::Config m_config;
    int m_cable;
    double m_startLength;
};

class tgKinematicActuatorInfo : public tgConnectorInfo {
public:
    explicit tgKinematicActuatorInfo(const tgKinematicActuator::Config& config);
    tgKinematicActuatorInfo(const tgKinematicActuator::Config& config, const tgTags& tags, const tgPair& pair);
    tgConnectorInfo* createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const;
    void initConnector(tgWorldB200Impl& world);
    tgModel* createModel(tgWorldB200Impl& world);
private:
    tgKinematicActuator::Config m_config;
    int m_cable;
    double m_startLength;
};

/* tag search -> info factory table (tgBuildSpec.cpp:58-74); later entries win. Owns the factories. */
class tgBuildSpec {
public:
    tgBuildSpec() {}
    ~tgBuildSpec();
    void addBuilder(std::string tag_search, tgRigidInfo* infoFactory);
    void addBuilder(std::string tag_search, tgConnectorInfo* infoFactory);
    struct RigidAgent { tgTagSearch tagSearch; tgRigidInfo* infoFactory; };
    struct ConnectorAgent { tgTagSearch tagSearch; tgConnectorInfo* infoFactory; };
    const std::vector<RigidAgent>& getRigidAgents() const { return m_rigidAgents; }
    const std::vector<ConnectorAgent>& getConnectorAgents() const { return m_connectorAgents; }
private:
    tgBuildSpec(const tgBuildSpec&);
    tgBuildSpec& operator=(const tgBuildSpec&);
    std::vector<RigidAgent> m_rigidAgents;
    std::vector<ConnectorAgent> m_connectorAgents;
};

/* structure + spec -> bodies, cables and the tgModel tree (tgStructureInfo.cpp:121-355). */
class tgStructureInfo : public tgTaggable {
public:
    tgStructureInfo(const tgStructure& structure, const tgBuildSpec& buildSpec);
    tgStructureInfo(const tgStructure& structure, const tgBuildSpec& buildSpec, const tgTags& tags);
    ~tgStructureInfo();
    void buildInto(tgModel& model, tgWorld& world);
    const std::vector<tgRigidInfo*>& getRigids() const { return m_rigids; }
    const std::vector<tgConnectorInfo*>& getConnectors() const { return m_connectors; }
private:
    tgStructureInfo(const tgStructureInfo&);
    tgStructureInfo& operator=(const tgStructureInfo&);
    void createChildren();
    void addRigidsAndConnectors();
    std::vector<tgRigidInfo*> getAllRigids() const;
    void autoCompoundRigids();
    void chooseConnectorRigids(const std::vector<tgRigidInfo*>& all);
    void initRigidBodies(tgWorldB200Impl& world);
    void initConnectors(tgWorldB200Impl& world);
    static void buildIntoHelper(tgModel& model, tgWorldB200Impl& world, tgStructureInfo& info);
    const tgStructure& m_structure;
    const tgBuildSpec& m_buildSpec;
    std::vector<tgRigidInfo*> m_rigids;
    std::vector<tgConnectorInfo*> m_connectors;
    std::vector<tgStructureInfo*> m_children;
    std::vector<tgRigidInfo*> m_compounds;
};

/* the reference's model files print their structure info while they build (tgStructureInfo.h:141-147, tgStructure.h:211-273) */
std::ostream& operator<<(std::ostream& os, const tgStructureInfo& info);
std::ostream& operator<<(std::ostream& os, const tgStructure& structure);

#endif /* TGB_CREATOR_H */
