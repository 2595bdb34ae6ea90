/*
 * tgb_creator.cpp -- structure -> flat descriptors. Reference arithmetic is followed where it decides
 * numbers (rod transform, mass through float, anchor relocation to the rod surface, rest length from
 * pretension); each such place cites the reference lines. Data flow and ownership are this repo's own.
 */
#include "tgb_creator.h"

#include <cstdlib>
#include <ctime>

#include <algorithm>
#include <cmath>
#include <sstream>
#include <queue>

/* ---- tgUtil ---------------------------------------------------------------------------------------- */
bool tgUtil::almostEqual(const btVector3& a, const btVector3& b, int precision) {
    return a.distance(b) < std::pow(10.0, (precision < 0 ? 5 : -precision));
}

static btVector3 arbitraryNonParallel(const btVector3& v) {
    /* any vector not parallel to v: swap in the axis with the smallest component */
    btVector3 a = v.absolute();
    if (a.x() <= a.y() && a.x() <= a.z()) return btVector3(1, 0, 0);
    if (a.y() <= a.z()) return btVector3(0, 1, 0);
    return btVector3(0, 0, 1);
}

/* tgUtil.h:199-238 */
btQuaternion tgUtil::getQuaternionBetween(btVector3 a, btVector3 b, const btVector3& fallbackAxis) {
    a.normalize();
    b.normalize();
    btQuaternion result;
    if (almostEqual(a, b)) {
        result = btQuaternion::getIdentity();
    } else if (almostEqual(a, -b)) {
        if (!fallbackAxis.isZero()) {
            result = btQuaternion(fallbackAxis, M_PI).normalize();
        } else if (a.dot(btVector3(1.0, 0.0, 0.0)) == 0.0) {
            result = btQuaternion(-1.0, 0.0, 0.0, 0.0);
        } else {
            const btVector3 arb = a + arbitraryNonParallel(a);
            const btVector3 c = (a.cross(arb)).normalize();
            result = btQuaternion(c, M_PI).normalize();
        }
    } else {
        const btVector3 c = (a.cross(b)).normalize();
        result = btQuaternion(c, std::acos(a.dot(b))).normalize();
    }
    return result;
}

/* tgUtil.h:92-107: origin at the midpoint, local Y turned onto (end - start) */
btTransform tgUtil::getTransform(const btVector3& start, const btVector3& end) {
    btTransform t;
    t.setIdentity();
    t.setOrigin(center(start, end));
    t.setRotation(getQuaternionBetween(upVector(), end - start));
    return t;
}

/* tgUtil.h:274-311 */
void tgUtil::addRotation(btVector3& v, const btVector3& fixedPoint, const btVector3& axis, double angle) {
    btVector3 relative = v - fixedPoint;
    btVector3 rotated = relative.rotate(axis, angle);
    v.setX(fixedPoint.x() + rotated.x());
    v.setY(fixedPoint.y() + rotated.y());
    v.setZ(fixedPoint.z() + rotated.z());
}
void tgUtil::addRotation(btVector3& v, const btVector3& fixedPoint, const btQuaternion& rotation) {
    addRotation(v, fixedPoint, rotation.getAxis(), rotation.getAngle());
}

/* ---- nodes / pairs / structure ----------------------------------------------------------------------- */
void tgPair::addRotation(const btVector3& fixedPoint, const btQuaternion& rotation) {
    tgUtil::addRotation(m_from, fixedPoint, rotation);
    tgUtil::addRotation(m_to, fixedPoint, rotation);
}
static void scalePoint(btVector3& p, const btVector3& ref, double f) {
    p.setX((p.x() - ref.x()) * f + ref.x());
    p.setY((p.y() - ref.y()) * f + ref.y());
    p.setZ((p.z() - ref.z()) * f + ref.z());
}
void tgPair::scale(const btVector3& ref, double f) { scalePoint(m_from, ref, f); scalePoint(m_to, ref, f); }

tgNode& tgNodes::operator[](std::size_t i) {
    if (i >= m_nodes.size()) throw std::out_of_range("tgNodes: index out of range");
    return m_nodes[i];
}
const tgNode& tgNodes::operator[](std::size_t i) const {
    if (i >= m_nodes.size()) throw std::out_of_range("tgNodes: index out of range");
    return m_nodes[i];
}
void tgNodes::move(const btVector3& offset) {
    for (std::size_t i = 0; i < m_nodes.size(); ++i) static_cast<btVector3&>(m_nodes[i]) += offset;
}
void tgNodes::addRotation(const btVector3& fixedPoint, const btQuaternion& rotation) {
    for (std::size_t i = 0; i < m_nodes.size(); ++i) tgUtil::addRotation(m_nodes[i], fixedPoint, rotation);
}
void tgNodes::scale(const btVector3& ref, double f) {
    for (std::size_t i = 0; i < m_nodes.size(); ++i) scalePoint(m_nodes[i], ref, f);
}
void tgPairs::move(const btVector3& offset) {
    for (std::size_t i = 0; i < m_pairs.size(); ++i) m_pairs[i].move(offset);
}
void tgPairs::addRotation(const btVector3& fixedPoint, const btQuaternion& rotation) {
    for (std::size_t i = 0; i < m_pairs.size(); ++i) m_pairs[i].addRotation(fixedPoint, rotation);
}
void tgPairs::scale(const btVector3& ref, double f) {
    for (std::size_t i = 0; i < m_pairs.size(); ++i) m_pairs[i].scale(ref, f);
}

tgStructure::tgStructure() {}
tgStructure::tgStructure(const tgTags& tags) : tgTaggable(tags) {}
tgStructure::tgStructure(const std::string& tags) : tgTaggable(tgTags(tags)) {}
tgStructure::tgStructure(const tgStructure& o) : tgTaggable(o.getTags()), m_nodes(o.m_nodes), m_pairs(o.m_pairs) {
    for (std::size_t i = 0; i < o.m_children.size(); ++i) m_children.push_back(new tgStructure(*o.m_children[i]));
}
tgStructure& tgStructure::operator=(const tgStructure& o) {
    if (this == &o) return *this;
    for (std::size_t i = 0; i < m_children.size(); ++i) delete m_children[i];
    m_children.clear();
    setTags(o.getTags());
    m_nodes = o.m_nodes;
    m_pairs = o.m_pairs;
    for (std::size_t i = 0; i < o.m_children.size(); ++i) m_children.push_back(new tgStructure(*o.m_children[i]));
    return *this;
}
tgStructure::~tgStructure() {
    for (std::size_t i = 0; i < m_children.size(); ++i) delete m_children[i];
}
void tgStructure::addNode(double x, double y, double z, std::string tags) { m_nodes.addNode(x, y, z, tags); }
void tgStructure::addNode(tgNode& n) { m_nodes.addNode(n); }
void tgStructure::addPair(int from, int to, std::string tags) {
    if (from < 0 || to < 0) throw std::out_of_range("tgStructure::addPair: negative node index");
    addPair(m_nodes[(std::size_t)from], m_nodes[(std::size_t)to], tags);
}
void tgStructure::addPair(const btVector3& from, const btVector3& to, std::string tags) {
    /* The reference's duplicate check compares addresses and therefore never fires (tgTaggables.h:106-114); shipped
       models rely on that (T6Model adds (2,4) and (4,2)), so duplicates are accepted here too. */
    m_pairs.addPair(tgPair(from, to, tags));
}
void tgStructure::move(const btVector3& offset) {
    m_nodes.move(offset);
    m_pairs.move(offset);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->move(offset);
}
void tgStructure::addRotation(const btVector3& fixedPoint, const btVector3& axis, double angle) {
    addRotation(fixedPoint, btQuaternion(axis, angle));
}
void tgStructure::addRotation(const btVector3& fixedPoint, const btVector3& from, const btVector3& to) {
    addRotation(fixedPoint, tgUtil::getQuaternionBetween(from, to));
}
void tgStructure::addRotation(const btVector3& fixedPoint, const btQuaternion& rotation) {
    m_nodes.addRotation(fixedPoint, rotation);
    m_pairs.addRotation(fixedPoint, rotation);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->addRotation(fixedPoint, rotation);
}
void tgStructure::scale(double f) { scale(getCentroid(), f); }
void tgStructure::scale(const btVector3& ref, double f) {
    m_nodes.scale(ref, f);
    m_pairs.scale(ref, f);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->scale(ref, f);
}
void tgStructure::addChild(tgStructure* child) {
    if (child) m_children.push_back(child);
}
void tgStructure::addChild(const tgStructure& child) { m_children.push_back(new tgStructure(child)); }
tgNode& tgStructure::findNode(const std::string& tags) {
    std::vector<tgStructure*> q(1, this);
    for (std::size_t head = 0; head < q.size(); ++head) {
        tgStructure* s = q[head];
        for (std::size_t i = 0; i < s->m_nodes.size(); ++i)
            if (s->m_nodes[i].hasAllTags(tags)) return s->m_nodes[i];
        q.insert(q.end(), s->m_children.begin(), s->m_children.end());
    }
    throw std::invalid_argument("Node not found: " + tags);
}
tgStructure& tgStructure::findChild(const std::string& tags) {
    std::vector<tgStructure*> q(m_children);
    for (std::size_t head = 0; head < q.size(); ++head) {
        tgStructure* s = q[head];
        if (s->hasAllTags(tags)) return *s;
        q.insert(q.end(), s->m_children.begin(), s->m_children.end());
    }
    throw std::invalid_argument("Child structure not found: " + tags);
}
tgPair& tgStructure::findPair(const btVector3& from, const btVector3& to) {
    std::vector<tgStructure*> q(1, this);
    for (std::size_t head = 0; head < q.size(); ++head) {
        tgStructure* s = q[head];
        for (std::size_t i = 0; i < s->m_pairs.size(); ++i) {
            const tgPair& p = s->m_pairs[i];
            if ((p.getFrom() == from && p.getTo() == to) || (p.getFrom() == to && p.getTo() == from)) return s->m_pairs[i];
        }
        q.insert(q.end(), s->m_children.begin(), s->m_children.end());
    }
    std::ostringstream os;
    os << "Pair not found: (" << from.x() << ", " << from.y() << ", " << from.z() << "), (" << to.x() << ", " << to.y() << ", " << to.z() << ")";
    throw std::invalid_argument(os.str());
}
void tgPairs::removePair(const tgPair& p) {
    const btVector3 from = p.getFrom(), to = p.getTo(); /* p may live in this very vector */
    std::size_t w = 0;
    for (std::size_t i = 0; i < m_pairs.size(); ++i)
        if (!(m_pairs[i].getFrom() == from && m_pairs[i].getTo() == to)) {
            if (w != i) m_pairs[w] = m_pairs[i];
            ++w;
        }
    m_pairs.resize(w);
}
void tgStructure::removePair(const tgPair& pair) {
    const tgPair copy(pair);
    m_pairs.removePair(copy);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->removePair(copy);
}
btVector3 tgStructure::getCentroid() const {
    btVector3 c(0, 0, 0);
    int n = 0;
    std::queue<const tgStructure*> q;
    q.push(this);
    while (!q.empty()) {
        const tgStructure* s = q.front();
        q.pop();
        for (std::size_t i = 0; i < s->m_nodes.size(); ++i) { c += s->m_nodes[i]; ++n; }
        for (std::size_t i = 0; i < s->m_children.size(); ++i) q.push(s->m_children[i]);
    }
    return c / n;
}

/* ---- rigid infos ------------------------------------------------------------------------------------- */
bool tgRigidInfo::sharesNodesWith(const tgRigidInfo& other) const {
    const std::vector<btVector3> a = getContainedNodes(), b = other.getContainedNodes();
    for (std::size_t i = 0; i < a.size(); ++i)
        for (std::size_t j = 0; j < b.size(); ++j)
            if (a[i] == b[j]) return true;
    return false;
}

/* One btRigidBody per group (tgRigidInfo.cpp:94-122): the first member to be initialised creates it. */
void tgRigidInfo::initRigidBody(tgWorldB200Impl& world) {
    if (getBodyIndex() >= 0) return;
    tgRigidInfo* rigid = getRigidInfoGroup() ? getRigidInfoGroup() : this;
    if (rigid->m_body >= 0) return;
    std::vector<tgb_shape_desc> shapes;
    tgb_body_desc body = rigid->describeBody(shapes);
    rigid->m_body = world.addBody(body, shapes);
}

static void fillCommonBody(tgb_body_desc& b, double mass, const btTransform& xf) {
    /* createRigidBody takes `float mass` (tgBulletUtil.cpp:39-42) */
    b.mass = (double)(float)mass;
    btVector3 o = xf.getOrigin();
    btQuaternion q = xf.getRotation();
    b.pos[0] = o.x(); b.pos[1] = o.y(); b.pos[2] = o.z();
    b.quat[0] = q.x(); b.quat[1] = q.y(); b.quat[2] = q.z(); b.quat[3] = q.w();
    b.gyroscopic = 1; /* tgRigidInfo.cpp:118 */
}

static double safeMargin(double radius, double halfLength) {
    /* btConvexInternalShape::setSafeMargin(halfExtents, 0.1) against CONVEX_DISTANCE_MARGIN 0.04 */
    double m = 0.1 * std::min(radius, halfLength);
    return m < 0.04 ? m : 0.04;
}

tgRodInfo::tgRodInfo(const tgRod::Config& config) : m_config(config) {}
tgRodInfo::tgRodInfo(const tgRod::Config& config, const tgTags& tags, const tgPair& pair)
    : tgRigidInfo(tags + pair.getTags()), m_pair(pair), m_config(config) {}
tgRigidInfo* tgRodInfo::createRigidInfo(const tgPair& pair, const tgTagSearch& search) const {
    if (!search.matches(pair.getTags())) return 0;
    return new tgRodInfo(m_config, getTags(), pair);
}
double tgRodInfo::getMass() const {
    /* tgRodInfo.cpp:116-123 */
    const double length = getLength();
    const double radius = m_config.radius;
    const double volume = M_PI * radius * radius * length;
    return volume * m_config.density;
}
std::vector<btVector3> tgRodInfo::getContainedNodes() const {
    std::vector<btVector3> v;
    v.push_back(getFrom());
    v.push_back(getTo());
    return v;
}
/* tgRodInfo.cpp:132-177: from the node, go `radius` sideways toward the other end of the cable */
btVector3 tgRodInfo::getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const {
    if (referencePoint == destinationPoint) throw std::invalid_argument("Destination point is the reference point.");
    const btVector3 cylinderAxis = (getTo() - getFrom()).normalize();
    const btVector3 refToDest = (referencePoint - destinationPoint).normalize();
    btVector3 rotationAxis = cylinderAxis.cross(refToDest);
    if (rotationAxis.length() == 0.0) {
        btScalar a = cylinderAxis[0], b = cylinderAxis[1], c = cylinderAxis[2];
        rotationAxis = btVector3(b - c, -a, a).normalize();
    }
    const btVector3 directional = cylinderAxis.rotate(rotationAxis, -M_PI / 2.0).normalize();
    const btVector3 surfaceVector = directional.rotate(cylinderAxis, rotation).normalize() * m_config.radius;
    return referencePoint + surfaceVector;
}
void tgRodInfo::appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const {
    tgb_shape_desc s = tgb_shape_desc();
    s.radius = m_config.radius;
    s.half_length = getLength() / 2.0;
    s.margin = safeMargin(s.radius, s.half_length);
    btTransform t = getTransform();
    btVector3 o = t.getOrigin() - bodyOrigin;
    btQuaternion q = t.getRotation();
    s.pos[0] = o.x(); s.pos[1] = o.y(); s.pos[2] = o.z();
    s.quat[0] = q.x(); s.quat[1] = q.y(); s.quat[2] = q.z(); s.quat[3] = q.w();
    out.push_back(s);
}
tgb_body_desc tgRodInfo::describeBody(std::vector<tgb_shape_desc>& shapes) const {
    tgb_body_desc b = tgb_body_desc();
    fillCommonBody(b, getMass(), getTransform());
    /* single shape sits at the body origin with identity rotation */
    tgb_shape_desc s = tgb_shape_desc();
    s.radius = m_config.radius;
    s.half_length = getLength() / 2.0;
    s.margin = safeMargin(s.radius, s.half_length);
    s.quat[3] = 1.0;
    shapes.push_back(s);
    if (b.mass != 0.0) {
        /* btCylinderShape::calculateLocalInertia, axis = local Y, half extents with margin */
        const double hr = (s.radius - s.margin) + s.margin, hh = (s.half_length - s.margin) + s.margin;
        const double radius2 = hr * hr, height2 = 4.0 * hh * hh;
        const double t1 = b.mass / 12.0 * height2 + b.mass / 4.0 * radius2;
        const double t2 = b.mass / 2.0 * radius2;
        b.inertia[0] = t1; b.inertia[1] = t2; b.inertia[2] = t1;
    }
    b.friction = m_config.friction;
    b.rolling_friction = m_config.rollFriction;
    b.restitution = m_config.restitution;
    return b;
}
tgModel* tgRodInfo::createModel(tgWorldB200Impl& world) {
    initRigidBody(world);
    const int body = getBodyIndex();
    return new tgRod(&world, body, world.bodies().at(body).mass, getTags(), getLength());
}

/* ---- box between two nodes ------------------------------------------------------------------------------- */
tgBoxInfo::tgBoxInfo(const tgBox::Config& config) : m_config(config) {}
tgBoxInfo::tgBoxInfo(const tgBox::Config& config, const tgTags& tags, const tgPair& pair)
    : tgRigidInfo(tags + pair.getTags()), m_pair(pair), m_config(config) {}
tgRigidInfo* tgBoxInfo::createRigidInfo(const tgPair& pair, const tgTagSearch& search) const {
    if (!search.matches(pair.getTags())) return 0;
    return new tgBoxInfo(m_config, getTags(), pair);
}
double tgBoxInfo::getMass() const {
    /* tgBoxInfo.cpp:120-133: width and height are half extents, the volume is that of the full box */
    const double volume = getLength() * (m_config.width * 2) * (m_config.height * 2);
    return volume * m_config.density;
}
std::vector<btVector3> tgBoxInfo::getContainedNodes() const {
    std::vector<btVector3> v;
    v.push_back(getFrom());
    v.push_back(getTo());
    return v;
}
/* tgBoxInfo.cpp:142-185: like a rod of radius `width` */
btVector3 tgBoxInfo::getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double rotation) const {
    if (referencePoint == destinationPoint) throw std::invalid_argument("Destination point is the reference point.");
    const btVector3 axis = (getTo() - getFrom()).normalize();
    const btVector3 refToDest = (referencePoint - destinationPoint).normalize();
    btVector3 rotationAxis = axis.cross(refToDest);
    if (rotationAxis.length() == 0.0) {
        btScalar a = axis[0], b = axis[1], c = axis[2];
        rotationAxis = btVector3(b - c, -a, a).normalize();
    }
    const btVector3 directional = axis.rotate(rotationAxis, -M_PI / 2.0).normalize();
    const btVector3 surfaceVector = directional.rotate(axis, rotation).normalize() * m_config.width;
    return referencePoint + surfaceVector;
}
tgb_shape_desc tgBoxInfo::shape() const {
    tgb_shape_desc s = tgb_shape_desc();
    s.kind = TGB_SHAPE_BOX;
    s.radius = m_config.width;          /* half extent along local X */
    s.half_length = getLength() / 2.0;  /* along local Y = the pair's axis */
    s.half_extent_z = m_config.height;
    /* btBoxShape: setSafeMargin(boxHalfExtents), btBoxShape.cpp / btConvexInternalShape.h */
    const double m = 0.1 * std::min(s.radius, std::min(s.half_length, s.half_extent_z));
    s.margin = m < 0.04 ? m : 0.04;
    s.quat[3] = 1.0;
    return s;
}
void tgBoxInfo::appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const {
    tgb_shape_desc s = shape();
    btTransform t = getTransform();
    btVector3 o = t.getOrigin() - bodyOrigin;
    btQuaternion q = t.getRotation();
    s.pos[0] = o.x(); s.pos[1] = o.y(); s.pos[2] = o.z();
    s.quat[0] = q.x(); s.quat[1] = q.y(); s.quat[2] = q.z(); s.quat[3] = q.w();
    out.push_back(s);
}
tgb_body_desc tgBoxInfo::describeBody(std::vector<tgb_shape_desc>& shapes) const {
    tgb_body_desc b = tgb_body_desc();
    fillCommonBody(b, getMass(), getTransform());
    const tgb_shape_desc s = shape();
    shapes.push_back(s);
    if (b.mass != 0.0) {
        /* btBoxShape::calculateLocalInertia: the box of the half extents with margin */
        const double lx = 2.0 * ((s.radius - s.margin) + s.margin), ly = 2.0 * ((s.half_length - s.margin) + s.margin),
                     lz = 2.0 * ((s.half_extent_z - s.margin) + s.margin);
        b.inertia[0] = b.mass / 12.0 * (ly * ly + lz * lz);
        b.inertia[1] = b.mass / 12.0 * (lx * lx + lz * lz);
        b.inertia[2] = b.mass / 12.0 * (lx * lx + ly * ly);
    }
    b.friction = m_config.friction;
    b.rolling_friction = m_config.rollFriction;
    b.restitution = m_config.restitution;
    return b;
}
tgModel* tgBoxInfo::createModel(tgWorldB200Impl& world) {
    initRigidBody(world);
    const int body = getBodyIndex();
    return new tgBox(&world, body, world.bodies().at(body).mass, getTags(), getLength());
}

/* ---- sphere at a node ------------------------------------------------------------------------------------ */
tgSphereInfo::tgSphereInfo(const tgSphere::Config& config) : m_config(config) {}
tgSphereInfo::tgSphereInfo(const tgSphere::Config& config, const tgTags& tags, const tgNode& node)
    : tgRigidInfo(tags + node.getTags()), m_node(node), m_config(config) {}
tgRigidInfo* tgSphereInfo::createRigidInfo(const tgNode& node, const tgTagSearch& search) const {
    if (!search.matches(node.getTags())) return 0;
    return new tgSphereInfo(m_config, getTags(), node);
}
double tgSphereInfo::getMass() const {
    /* tgSphereInfo.cpp:117-123 */
    const double radius = m_config.radius;
    const double volume = 4.0 / 3.0 * M_PI * radius * radius * radius;
    return volume * m_config.density;
}
btTransform tgSphereInfo::getTransform() const {
    /* tgSphereInfo.h: identity rotation at the node */
    btTransform t;
    t.setIdentity();
    t.setOrigin(m_node);
    return t;
}
/* tgSphereInfo.cpp:133-163: the surface point toward the destination */
btVector3 tgSphereInfo::getConnectionPoint(const btVector3& referencePoint, const btVector3& destinationPoint, double) const {
    if (referencePoint == destinationPoint) throw std::invalid_argument("Destination point is the reference point.");
    const btVector3 refToDest = (destinationPoint - referencePoint).normalize();
    return m_node + refToDest * m_config.radius;
}
void tgSphereInfo::appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const {
    tgb_shape_desc s = tgb_shape_desc();
    s.kind = TGB_SHAPE_SPHERE;
    s.radius = m_config.radius;
    s.half_length = m_config.radius; /* btSphereShape: a point with collision margin = radius */
    s.margin = m_config.radius;
    const btVector3 o = m_node - bodyOrigin;
    s.pos[0] = o.x(); s.pos[1] = o.y(); s.pos[2] = o.z();
    s.quat[3] = 1.0;
    out.push_back(s);
}
tgb_body_desc tgSphereInfo::describeBody(std::vector<tgb_shape_desc>& shapes) const {
    tgb_body_desc b = tgb_body_desc();
    fillCommonBody(b, getMass(), getTransform());
    appendShapes(m_node, shapes);
    if (b.mass != 0.0) {
        /* btSphereShape::calculateLocalInertia: 0.4 m r^2 on every axis */
        const double e = 0.4 * b.mass * m_config.radius * m_config.radius;
        b.inertia[0] = e; b.inertia[1] = e; b.inertia[2] = e;
    }
    b.friction = m_config.friction;
    b.rolling_friction = m_config.rollFriction;
    b.restitution = m_config.restitution;
    return b;
}
tgModel* tgSphereInfo::createModel(tgWorldB200Impl& world) {
    initRigidBody(world);
    const int body = getBodyIndex();
    return new tgSphere(&world, body, world.bodies().at(body).mass, getTags());
}

double tgCompoundRigidInfo::getMass() const {
    double m = 0.0;
    for (std::size_t i = 0; i < m_rigids.size(); ++i) m += m_rigids[i]->getMass();
    return m;
}
btVector3 tgCompoundRigidInfo::getCenterOfMass() const {
    btVector3 sum(0, 0, 0);
    for (std::size_t i = 0; i < m_rigids.size(); ++i) sum += m_rigids[i]->getCenterOfMass() * m_rigids[i]->getMass();
    const double total = getMass();
    return total == 0.0 ? btVector3(0, 0, 0) : sum / total;
}
btTransform tgCompoundRigidInfo::getTransform() const {
    btTransform t;
    t.setIdentity();
    t.setOrigin(getCenterOfMass());
    return t;
}
bool tgCompoundRigidInfo::containsNode(const btVector3& node) const {
    for (std::size_t i = 0; i < m_rigids.size(); ++i)
        if (m_rigids[i]->containsNode(node)) return true;
    return false;
}
std::vector<btVector3> tgCompoundRigidInfo::getContainedNodes() const {
    std::vector<btVector3> all;
    for (std::size_t i = 0; i < m_rigids.size(); ++i) {
        std::vector<btVector3> n = m_rigids[i]->getContainedNodes();
        all.insert(all.end(), n.begin(), n.end());
    }
    return all;
}
btVector3 tgCompoundRigidInfo::getConnectionPoint(const btVector3& ref, const btVector3& dest, double rotation) const {
    for (std::size_t i = 0; i < m_rigids.size(); ++i)
        if (m_rigids[i]->containsNode(ref)) return m_rigids[i]->getConnectionPoint(ref, dest, rotation);
    return ref;
}
void tgCompoundRigidInfo::appendShapes(const btVector3& bodyOrigin, std::vector<tgb_shape_desc>& out) const {
    for (std::size_t i = 0; i < m_rigids.size(); ++i) m_rigids[i]->appendShapes(bodyOrigin, out);
}
tgb_body_desc tgCompoundRigidInfo::describeBody(std::vector<tgb_shape_desc>& shapes) const {
    tgb_body_desc b = tgb_body_desc();
    const btVector3 com = getCenterOfMass();
    fillCommonBody(b, getMass(), getTransform());
    appendShapes(com, shapes);
    /* btCompoundShape::calculateLocalInertia = box inertia of the local AABB of the children, each child's AABB being
       that of its margin-inclusive half extents under its child transform (SURVEY Appendix B "Compound inertia") */
    btVector3 lo(1e300, 1e300, 1e300), hi(-1e300, -1e300, -1e300);
    for (std::size_t i = 0; i < shapes.size(); ++i) {
        const tgb_shape_desc& s = shapes[i];
        btMatrix3x3 R(btQuaternion(s.quat[0], s.quat[1], s.quat[2], s.quat[3]));
        btMatrix3x3 A = R.absolute();
        const double hz = s.kind == TGB_SHAPE_BOX ? s.half_extent_z : s.radius;
        btVector3 he((s.radius - s.margin) + s.margin, (s.half_length - s.margin) + s.margin, (hz - s.margin) + s.margin);
        btVector3 ext = he.dot3(A[0], A[1], A[2]);
        btVector3 c(s.pos[0], s.pos[1], s.pos[2]);
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], c[k] - ext[k]);
            hi[k] = std::max(hi[k], c[k] + ext[k]);
        }
    }
    if (b.mass != 0.0 && !shapes.empty()) {
        btVector3 half = (hi - lo) * 0.5;
        const double lx = 2.0 * half.x(), ly = 2.0 * half.y(), lz = 2.0 * half.z();
        b.inertia[0] = b.mass / 12.0 * (ly * ly + lz * lz);
        b.inertia[1] = b.mass / 12.0 * (lx * lx + lz * lz);
        b.inertia[2] = b.mass / 12.0 * (lx * lx + ly * ly);
    }
    getMaterial(b.friction, b.rolling_friction, b.restitution);
    return b;
}
/* Every leaf's initRigidBody sets friction / rolling friction / restitution on the shared btRigidBody after creating it
   (tgRodInfo.cpp:72-79, tgSphereInfo.cpp:71-78), in the order the structure info initialises its rigids, so the last member's
   Config is the one that stays. */
void tgCompoundRigidInfo::getMaterial(double& friction, double& rollFriction, double& restitution) const {
    friction = 0.5; rollFriction = 0.0; restitution = 0.0;
    if (!m_rigids.empty()) m_rigids.back()->getMaterial(friction, rollFriction, restitution);
}
tgModel* tgCompoundRigidInfo::createModel(tgWorldB200Impl&) { return 0; } /* leaves create their own tgRod wrappers */

/* ---- connectors ---------------------------------------------------------------------------------------- */
void tgConnectorInfo::chooseRigids(const std::vector<tgRigidInfo*>& rigids) {
    /* tgConnectorInfo.cpp:62-117: the rigid that contains the node; several -> closest centre of mass */
    struct Pick {
        static tgRigidInfo* best(const std::vector<tgRigidInfo*>& all, const btVector3& v) {
            tgRigidInfo* chosen = 0;
            for (std::size_t i = 0; i < all.size(); ++i) {
                if (!all[i]->containsNode(v)) continue;
                if (!chosen || all[i]->getCenterOfMass().distance(v) < chosen->getCenterOfMass().distance(v)) chosen = all[i];
            }
            return chosen;
        }
    };
    if (!m_fromRigid) m_fromRigid = Pick::best(rigids, getFrom());
    if (!m_toRigid) m_toRigid = Pick::best(rigids, getTo());
}

tgb_cable_desc tgConnectorInfo::describeCable(const tgSpringCableActuator::Config& cfg, tgWorldB200Impl& world, double& startLength) const {
    if (!m_fromRigid || !m_toRigid) throw std::runtime_error("getFromRigidInfo == NULL, cannot return rigid body.");
    /* tgBasicActuatorInfo.cpp:77-125: optionally move each end from the node to the rod surface */
    btVector3 from = cfg.moveCablePointAToEdge ? m_fromRigid->getConnectionPoint(getFrom(), getTo(), cfg.rotation) : getFrom();
    btVector3 to = cfg.moveCablePointBToEdge ? m_toRigid->getConnectionPoint(getTo(), getFrom(), cfg.rotation) : getTo();
    tgb_cable_desc c = tgb_cable_desc();
    c.body_a = m_fromRigid->getBodyIndex();
    c.body_b = m_toRigid->getBodyIndex();
    if (c.body_a < 0 || c.body_b < 0) throw std::runtime_error("connector end has no rigid body");
    const tgb_body_desc& A = world.bodies().at(c.body_a);
    const tgb_body_desc& B = world.bodies().at(c.body_b);
    /* anchor local point = bodyTransform^-1 * worldPoint (tgBulletSpringCableAnchor.cpp:54-55) */
    btTransform ta(btQuaternion(A.quat[0], A.quat[1], A.quat[2], A.quat[3]), btVector3(A.pos[0], A.pos[1], A.pos[2]));
    btTransform tb(btQuaternion(B.quat[0], B.quat[1], B.quat[2], B.quat[3]), btVector3(B.pos[0], B.pos[1], B.pos[2]));
    btVector3 la = ta.inverse() * from, lb = tb.inverse() * to;
    for (int k = 0; k < 3; ++k) { c.local_a[k] = la[k]; c.local_b[k] = lb[k]; }
    c.stiffness = cfg.stiffness;
    c.damping = cfg.damping;
    /* tgSpringCable.cpp:49-59: rest = distance of the anchors' world positions - pretension / k; prev = rest */
    btVector3 pos1 = ta * la, pos2 = tb * lb;
    startLength = pos1.distance(pos2);
    c.rest_length = startLength - cfg.pretension / cfg.stiffness;
    if (c.rest_length <= 0.0) throw std::invalid_argument("Pretension causes string to shorten past rest length!");
    c.prev_length = c.rest_length;
    c.max_tension = cfg.maxTens;
    c.target_velocity = cfg.targetVelocity;
    c.min_actual_length = cfg.minActualLength;
    c.min_rest_length = cfg.minRestLength;
    c.motor_radius = 1.0; c.motor_friction = 0.0; c.motor_inertia = 1.0;
    return c;
}

tgBasicActuatorInfo::tgBasicActuatorInfo(const tgBasicActuator::Config& config) : m_config(config), m_cable(-1), m_startLength(0) {}
tgBasicActuatorInfo::tgBasicActuatorInfo(const tgBasicActuator::Config& config, const tgTags& tags, const tgPair& pair)
    : tgConnectorInfo(tags + pair.getTags(), pair), m_config(config), m_cable(-1), m_startLength(0) {}
tgConnectorInfo* tgBasicActuatorInfo::createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const {
    if (!search.matches(pair.getTags())) return 0;
    return new tgBasicActuatorInfo(m_config, getTags(), pair);
}
void tgBasicActuatorInfo::initConnector(tgWorldB200Impl& world) {
    tgb_cable_desc c = describeCable(m_config, world, m_startLength);
    c.actuator = TGB_ACT_BASIC;
    m_cable = world.addCable(c);
}
tgModel* tgBasicActuatorInfo::createModel(tgWorldB200Impl& world) {
    if (m_cable < 0) initConnector(world);
    return new tgBasicActuator(&world, m_cable, getTags(), m_config, m_startLength);
}

tgBasicContactCableInfo::tgBasicContactCableInfo(const tgBasicActuator::Config& config) : m_config(config), m_cable(-1), m_startLength(0) {}
tgBasicContactCableInfo::tgBasicContactCableInfo(const tgBasicActuator::Config& config, const tgTags& tags, const tgPair& pair)
    : tgConnectorInfo(tags + pair.getTags(), pair), m_config(config), m_cable(-1), m_startLength(0) {}
tgConnectorInfo* tgBasicContactCableInfo::createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const {
    if (!search.matches(pair.getTags())) return 0;
    return new tgBasicContactCableInfo(m_config, getTags(), pair);
}
void tgBasicContactCableInfo::initConnector(tgWorldB200Impl& world) {
    tgSpringCableActuator::Config cfg = m_config;
    cfg.moveCablePointAToEdge = cfg.moveCablePointBToEdge = true; /* tgBasicContactCableInfo.cpp:90-91 */
    const tgb_cable_desc plain = describeCable(cfg, world, m_startLength);
    tgb_contact_cable_desc c = tgb_contact_cable_desc();
    c.body_a = plain.body_a; c.body_b = plain.body_b;
    for (int k = 0; k < 3; ++k) { c.local_a[k] = plain.local_a[k]; c.local_b[k] = plain.local_b[k]; }
    c.stiffness = plain.stiffness; c.damping = plain.damping;
    c.rest_length = plain.rest_length; c.prev_length = plain.prev_length;
    c.thickness = 0.001; c.resolution = 0.1;
    c.max_tension = plain.max_tension; c.target_velocity = plain.target_velocity;
    c.min_actual_length = plain.min_actual_length; c.min_rest_length = plain.min_rest_length;
    m_cable = world.addContactCable(c);
}
tgModel* tgBasicContactCableInfo::createModel(tgWorldB200Impl& world) {
    if (m_cable < 0) initConnector(world);
    return new tgBasicActuator(&world, m_cable, getTags(), m_config, m_startLength);
}

tgKinematicActuatorInfo::tgKinematicActuatorInfo(const tgKinematicActuator::Config& config) : m_config(config), m_cable(-1), m_startLength(0) {}
tgKinematicActuatorInfo::tgKinematicActuatorInfo(const tgKinematicActuator::Config& config, const tgTags& tags, const tgPair& pair)
    : tgConnectorInfo(tags + pair.getTags(), pair), m_config(config), m_cable(-1), m_startLength(0) {}
tgConnectorInfo* tgKinematicActuatorInfo::createConnectorInfo(const tgPair& pair, const tgTagSearch& search) const {
    if (!search.matches(pair.getTags())) return 0;
    return new tgKinematicActuatorInfo(m_config, getTags(), pair);
}
void tgKinematicActuatorInfo::initConnector(tgWorldB200Impl& world) {
    tgb_cable_desc c = describeCable(m_config, world, m_startLength);
    c.actuator = TGB_ACT_KINEMATIC;
    c.backdrivable = m_config.backdrivable ? 1 : 0;
    c.motor_radius = m_config.radius;
    c.motor_friction = m_config.motorFriction;
    c.motor_inertia = m_config.motorInertia;
    m_cable = world.addCable(c);
}
tgModel* tgKinematicActuatorInfo::createModel(tgWorldB200Impl& world) {
    if (m_cable < 0) initConnector(world);
    return new tgKinematicActuator(&world, m_cable, getTags(), m_config, m_startLength);
}

/* ---- build spec ------------------------------------------------------------------------------------------ */
tgBuildSpec::~tgBuildSpec() {
    for (std::size_t i = 0; i < m_rigidAgents.size(); ++i) delete m_rigidAgents[i].infoFactory;
    for (std::size_t i = 0; i < m_connectorAgents.size(); ++i) delete m_connectorAgents[i].infoFactory;
}
void tgBuildSpec::addBuilder(std::string tag_search, tgRigidInfo* infoFactory) {
    RigidAgent a;
    a.tagSearch = tgTagSearch(tag_search);
    a.infoFactory = infoFactory;
    m_rigidAgents.push_back(a);
}
void tgBuildSpec::addBuilder(std::string tag_search, tgConnectorInfo* infoFactory) {
    ConnectorAgent a;
    a.tagSearch = tgTagSearch(tag_search);
    a.infoFactory = infoFactory;
    m_connectorAgents.push_back(a);
}

/* ---- structure info -------------------------------------------------------------------------------------- */
tgStructureInfo::tgStructureInfo(const tgStructure& structure, const tgBuildSpec& buildSpec) : m_structure(structure), m_buildSpec(buildSpec) {
    createChildren();
}
tgStructureInfo::tgStructureInfo(const tgStructure& structure, const tgBuildSpec& buildSpec, const tgTags& tags)
    : tgTaggable(tags), m_structure(structure), m_buildSpec(buildSpec) {
    createChildren();
}
tgStructureInfo::~tgStructureInfo() {
    for (std::size_t i = 0; i < m_compounds.size(); ++i) delete m_compounds[i];
    for (std::size_t i = 0; i < m_rigids.size(); ++i) delete m_rigids[i];
    for (std::size_t i = 0; i < m_connectors.size(); ++i) delete m_connectors[i];
    for (std::size_t i = 0; i < m_children.size(); ++i) delete m_children[i];
}
void tgStructureInfo::createChildren() {
    const std::vector<tgStructure*>& kids = m_structure.getChildren();
    for (std::size_t i = 0; i < kids.size(); ++i) m_children.push_back(new tgStructureInfo(*kids[i], m_buildSpec, kids[i]->getTags()));
}
/* tgStructureInfo.cpp:121-198: every pair becomes a rigid if a rigid agent matches (later agents first), else a
   connector if a connector agent matches. The structure's own tags are inherited by its members. */
void tgStructureInfo::addRigidsAndConnectors() {
    const std::vector<tgBuildSpec::RigidAgent>& ra = m_buildSpec.getRigidAgents();
    const std::vector<tgBuildSpec::ConnectorAgent>& ca = m_buildSpec.getConnectorAgents();
    const tgPairs& pairs = m_structure.getPairs();
    /* nodes first (tgStructureInfo.cpp:128-135): a tagged node becomes a rigid if a rigid agent answers for nodes */
    const tgNodes& nodes = m_structure.getNodes();
    for (std::size_t n = 0; n < nodes.size(); ++n) {
        tgRigidInfo* rigid = 0;
        for (int i = (int)ra.size() - 1; i >= 0 && !rigid; --i) {
            tgTagSearch s = ra[i].tagSearch;
            s.remove(getTags());
            rigid = ra[i].infoFactory->createRigidInfo(nodes[n], s);
        }
        if (rigid) m_rigids.push_back(rigid);
    }
    for (std::size_t p = 0; p < pairs.size(); ++p) {
        tgRigidInfo* rigid = 0;
        for (int i = (int)ra.size() - 1; i >= 0 && !rigid; --i) {
            tgTagSearch s = ra[i].tagSearch;
            s.remove(getTags());
            rigid = ra[i].infoFactory->createRigidInfo(pairs[p], s);
        }
        if (rigid) { m_rigids.push_back(rigid); continue; }
        tgConnectorInfo* conn = 0;
        for (int i = (int)ca.size() - 1; i >= 0 && !conn; --i) {
            tgTagSearch s = ca[i].tagSearch;
            s.remove(getTags());
            conn = ca[i].infoFactory->createConnectorInfo(pairs[p], s);
        }
        if (conn) m_connectors.push_back(conn);
    }
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->addRigidsAndConnectors();
}
std::vector<tgRigidInfo*> tgStructureInfo::getAllRigids() const {
    std::vector<tgRigidInfo*> all(m_rigids);
    for (std::size_t i = 0; i < m_children.size(); ++i) {
        std::vector<tgRigidInfo*> sub = m_children[i]->getAllRigids();
        all.insert(all.end(), sub.begin(), sub.end());
    }
    return all;
}
/* tgRigidAutoCompound.cpp:56-157: connected components of the "shares a node" graph, discovered depth-first from the
   first ungrouped rigid; a component of one keeps itself as its group. */
void tgStructureInfo::autoCompoundRigids() {
    std::vector<tgRigidInfo*> all = getAllRigids();
    std::vector<char> used(all.size(), 0);
    for (std::size_t seed = 0; seed < all.size(); ++seed) {
        if (used[seed]) continue;
        std::vector<std::size_t> group, stack;
        used[seed] = 1;
        stack.push_back(seed);
        while (!stack.empty()) {
            std::size_t cur = stack.back();
            stack.pop_back();
            group.push_back(cur);
            for (std::size_t j = 0; j < all.size(); ++j) {
                if (used[j] || !all[cur]->sharesNodesWith(*all[j])) continue;
                used[j] = 1;
                stack.push_back(j);
            }
        }
        if (group.size() == 1) {
            all[group[0]]->setRigidInfoGroup(all[group[0]]);
        } else {
            std::sort(group.begin(), group.end());
            tgCompoundRigidInfo* c = new tgCompoundRigidInfo();
            /* every member carries the tag "compound_" + six characters of [0-9A-Za-z] by which tgCompoundRigidSensorInfo finds the
               members of one compound (tgRigidAutoCompound.cpp:146-157; the reference draws the characters from a random device,
               here a counter-seeded generator: tags are unique per process and reproducible) */
            static unsigned long long compoundCounter = 0;
            unsigned long long h = (++compoundCounter) * 6364136223846793005ULL + 1442695040888963407ULL;
            static const char alphanum[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz";
            std::string compoundTag = "compound_";
            for (int k = 0; k < 6; ++k) {
                h = h * 6364136223846793005ULL + 1442695040888963407ULL;
                compoundTag += alphanum[(h >> 33) % 62];
            }
            for (std::size_t g = 0; g < group.size(); ++g) {
                all[group[g]]->addTags(compoundTag);
                c->addRigid(*all[group[g]]);
                all[group[g]]->setRigidInfoGroup(c);
            }
            c->setRigidInfoGroup(c);
            m_compounds.push_back(c);
        }
    }
}
void tgStructureInfo::chooseConnectorRigids(const std::vector<tgRigidInfo*>& all) {
    for (std::size_t i = 0; i < m_connectors.size(); ++i) m_connectors[i]->chooseRigids(all);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->chooseConnectorRigids(all);
}
void tgStructureInfo::initRigidBodies(tgWorldB200Impl& world) {
    for (std::size_t i = 0; i < m_rigids.size(); ++i) m_rigids[i]->initRigidBody(world);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->initRigidBodies(world);
}
void tgStructureInfo::initConnectors(tgWorldB200Impl& world) {
    for (std::size_t i = 0; i < m_connectors.size(); ++i) m_connectors[i]->initConnector(world);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->initConnectors(world);
}
/* tgStructureInfo.cpp:313-355: rigids, then connectors, then one sub-model per child structure */
void tgStructureInfo::buildIntoHelper(tgModel& model, tgWorldB200Impl& world, tgStructureInfo& info) {
    for (std::size_t i = 0; i < info.m_rigids.size(); ++i) {
        tgModel* m = info.m_rigids[i]->createModel(world);
        if (m) { m->setTags(info.m_rigids[i]->getTags()); model.addChild(m); }
    }
    for (std::size_t i = 0; i < info.m_connectors.size(); ++i) {
        tgModel* m = info.m_connectors[i]->createModel(world);
        if (m) { m->setTags(info.m_connectors[i]->getTags()); model.addChild(m); }
    }
    for (std::size_t i = 0; i < info.m_children.size(); ++i) {
        tgModel* sub = new tgModel();
        buildIntoHelper(*sub, world, *info.m_children[i]);
        model.addChild(sub);
    }
    model.setTags(info.getTags());
}
void tgStructureInfo::buildInto(tgModel& model, tgWorld& world) {
    tgWorldB200Impl& impl = world.b200();
    addRigidsAndConnectors();
    autoCompoundRigids();
    chooseConnectorRigids(getAllRigids());
    initRigidBodies(impl);
    initConnectors(impl);
    buildIntoHelper(model, impl, *this);
}

/* tgUtil.cpp:54-62 */
void tgUtil::seedRandom() { srand((unsigned)time(0)); }
void tgUtil::seedRandom(int seed) { srand((unsigned)seed); }

std::ostream& operator<<(std::ostream& os, const tgStructure& structure) {
    os << "tgStructure(" << structure.getNodes().size() << " nodes, " << structure.getPairs().size() << " pairs, " << structure.getChildren().size()
       << " children) tags: " << structure.getTags();
    return os;
}

std::ostream& operator<<(std::ostream& os, const tgStructureInfo& info) {
    os << "tgStructureInfo(" << info.getRigids().size() << " rigids, " << info.getConnectors().size() << " connectors)";
    return os;
}
