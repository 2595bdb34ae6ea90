/*
 * tgb_host.cpp -- implementation of the host-side core API declared in tgb_host.h.
 * Reference behaviour followed is cited per function (paths relative to the NTRTsim tree).
 */
#include "tgb_host.h"
#include "tgb_sensors.h"

#include <algorithm>
#include <cmath>
#include <iostream>

/* ---- tags ---------------------------------------------------------------------------------------- */
void tgTags::append(const std::string& s) {
    std::istringstream is(s);
    std::string word;
    while (is >> word)
        if (!contains(word)) m_words.push_back(word);
}
void tgTags::append(const tgTags& o) {
    for (std::size_t i = 0; i < o.m_words.size(); ++i)
        if (!contains(o.m_words[i])) m_words.push_back(o.m_words[i]);
}
void tgTags::remove(const tgTags& o) {
    for (std::size_t i = 0; i < o.m_words.size(); ++i)
        m_words.erase(std::remove(m_words.begin(), m_words.end(), o.m_words[i]), m_words.end());
}
bool tgTags::contains(const std::string& w) const { return std::find(m_words.begin(), m_words.end(), w) != m_words.end(); }
bool tgTags::containsAll(const tgTags& o) const {
    for (std::size_t i = 0; i < o.m_words.size(); ++i)
        if (!contains(o.m_words[i])) return false;
    return true;
}
bool tgTags::containsAny(const tgTags& o) const {
    for (std::size_t i = 0; i < o.m_words.size(); ++i)
        if (contains(o.m_words[i])) return true;
    return false;
}
std::string tgTags::joinTags(const std::string& sep) const {
    std::string out;
    for (std::size_t i = 0; i < m_words.size(); ++i) {
        if (i) out += sep;
        out += m_words[i];
    }
    return out;
}
std::ostream& operator<<(std::ostream& os, const tgTags& t) { return os << t.joinTags(" "); }

tgTagSearch::tgTagSearch(const std::string& expr) { parse(expr); }
void tgTagSearch::parse(const std::string& expr) {
    std::size_t start = 0;
    while (start <= expr.size()) {
        std::size_t bar = expr.find('|', start);
        std::string part = expr.substr(start, bar == std::string::npos ? std::string::npos : bar - start);
        Alternative alt;
        std::istringstream is(part);
        std::string word;
        while (is >> word) {
            if (word.size() > 1 && word[0] == '-') alt.forbidden.append(word.substr(1));
            else alt.required.append(word);
        }
        m_alternatives.push_back(alt);
        if (bar == std::string::npos) break;
        start = bar + 1;
    }
}
bool tgTagSearch::matches(const tgTags& tags) const {
    for (std::size_t i = 0; i < m_alternatives.size(); ++i) {
        const Alternative& a = m_alternatives[i];
        if (a.required.empty() && a.forbidden.empty()) continue;
        if (tags.containsAll(a.required) && !tags.containsAny(a.forbidden)) return true;
    }
    return false;
}
void tgTagSearch::remove(const tgTags& tags) {
    for (std::size_t i = 0; i < m_alternatives.size(); ++i) m_alternatives[i].required.remove(tags);
}

/* ---- grounds -------------------------------------------------------------------------------------- */
tgb_ground_desc tgEmptyGround::describe() const {
    tgb_ground_desc g = tgb_ground_desc();
    g.kind = TGB_GROUND_EMPTY;
    g.normal[1] = 1.0;
    return g;
}

tgBoxGround::Config::Config(btVector3 eulerAngles, btScalar friction, btScalar restitution, btVector3 size, btVector3 origin)
    : m_eulerAngles(eulerAngles), m_friction(friction), m_restitution(restitution), m_size(size), m_origin(origin) {
    if (friction < 0.0 || friction > 1.0) throw std::invalid_argument("tgBoxGround friction outside [0,1]");
    if (restitution < 0.0 || restitution > 1.0) throw std::invalid_argument("tgBoxGround restitution outside [0,1]");
    if (size[0] < 0.0 || size[1] < 0.0 || size[2] < 0.0) throw std::invalid_argument("tgBoxGround size negative");
}
tgBoxGround::tgBoxGround() : m_config(Config()) {}
tgBoxGround::tgBoxGround(const Config& config) : m_config(config) {}
tgb_ground_desc tgBoxGround::describe() const {
    /* body transform: origin + setEuler(yaw, pitch, roll) (tgBoxGround.cpp:77-86); contact surface = top face */
    btQuaternion q;
    q.setEuler(m_config.m_eulerAngles[0], m_config.m_eulerAngles[1], m_config.m_eulerAngles[2]);
    btMatrix3x3 R(q);
    btVector3 n = R * btVector3(0.0, 1.0, 0.0);
    btVector3 top = m_config.m_origin + n * m_config.m_size[1];
    tgb_ground_desc g = tgb_ground_desc();
    g.kind = TGB_GROUND_BOX;
    g.normal[0] = n.x(); g.normal[1] = n.y(); g.normal[2] = n.z();
    g.offset = n.dot(top);
    /* Config friction/restitution never reach the body in the reference (tgBoxGround.cpp:73-100): Bullet defaults */
    g.friction = 0.5;
    g.rolling_friction = 0.0;
    g.restitution = 0.0;
    double minHalf = std::min(m_config.m_size[0], std::min(m_config.m_size[1], m_config.m_size[2]));
    g.margin = std::min(0.04, 0.1 * minHalf); /* btConvexInternalShape::setSafeMargin */
    return g;
}

tgPlaneGround::Config::Config(btVector3 normalVector, btScalar friction, btScalar restitution, btVector3 origin)
    : m_normalVector(normalVector), m_friction(friction), m_restitution(restitution), m_origin(origin) {
    if (friction < 0.0 || friction > 1.0) throw std::invalid_argument("tgPlaneGround friction outside [0,1]");
    if (restitution < 0.0 || restitution > 1.0) throw std::invalid_argument("tgPlaneGround restitution outside [0,1]");
}
tgPlaneGround::tgPlaneGround() : m_config(Config()) {}
tgPlaneGround::tgPlaneGround(const Config& config) : m_config(config) {}
tgb_ground_desc tgPlaneGround::describe() const {
    btVector3 n = m_config.m_normalVector.normalized();
    tgb_ground_desc g = tgb_ground_desc();
    g.kind = TGB_GROUND_PLANE;
    g.normal[0] = n.x(); g.normal[1] = n.y(); g.normal[2] = n.z();
    g.offset = n.dot(m_config.m_origin); /* btStaticPlaneShape(n, 0) placed at origin (tgPlaneGround.cpp:56-75) */
    g.friction = m_config.m_friction;
    g.rolling_friction = 0.0;
    g.restitution = m_config.m_restitution;
    g.margin = 0.04;
    return g;
}

/* ---- batched world impl --------------------------------------------------------------------------- */
tgWorldB200Impl::tgWorldB200Impl(double gravity, const tgGround* ground)
    : m_gravity(gravity), m_nEnv(1), m_precision(TGB_F32), m_device(0), m_handle(0), m_cacheValid(false) {
    if (ground) m_ground = ground->describe();
    else m_ground = tgEmptyGround().describe();
}
tgWorldB200Impl::~tgWorldB200Impl() {
    if (m_handle) tgb_destroy(m_handle);
}
int tgWorldB200Impl::addBody(const tgb_body_desc& body, const std::vector<tgb_shape_desc>& shapes) {
    if (m_handle) throw std::runtime_error("tgWorldB200Impl: cannot add bodies after the world was finalized");
    tgb_body_desc b = body;
    b.shape_begin = (int32_t)m_shapes.size();
    b.shape_count = (int32_t)shapes.size();
    m_shapes.insert(m_shapes.end(), shapes.begin(), shapes.end());
    m_bodies.push_back(b);
    return (int)m_bodies.size() - 1;
}
int tgWorldB200Impl::addCable(const tgb_cable_desc& cable) {
    if (m_handle) throw std::runtime_error("tgWorldB200Impl: cannot add cables after the world was finalized");
    m_cables.push_back(cable);
    return (int)m_cables.size() - 1;
}
int tgWorldB200Impl::addContactCable(const tgb_contact_cable_desc& cable) {
    if (m_handle) throw std::runtime_error("tgWorldB200Impl: cannot add cables after the world was finalized");
    m_contactCables.push_back(cable);
    return kContactCableBase + (int)m_contactCables.size() - 1;
}
void tgWorldB200Impl::setBatch(int n_env, tgb_precision precision, int device) {
    if (n_env < 1) throw std::invalid_argument("n_env must be positive");
    if (m_handle) throw std::runtime_error("tgWorldB200Impl: batch size is fixed once the world was finalized");
    m_nEnv = n_env; m_precision = precision; m_device = device;
}
void tgWorldB200Impl::check(int status) const {
    if (status == TGB_OK) return;
    std::string msg = tgb_last_error();
    if (status == TGB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
void tgWorldB200Impl::finalize() {
    if (m_handle) return;
    tgb_world_cfg cfg = tgb_world_cfg();
    cfg.gravity = m_gravity; cfg.n_env = m_nEnv; cfg.precision = m_precision; cfg.device = m_device; cfg.solver_iterations = 0;
    tgb_world* h = 0;
    check(tgb_create(&cfg, &m_ground, &h));
    int st = tgb_add_model_flat(h, m_bodies.data(), (int32_t)m_bodies.size(), m_shapes.data(), (int32_t)m_shapes.size(), m_cables.data(),
                                (int32_t)m_cables.size());
    if (st == TGB_OK && !m_contactCables.empty()) st = tgb_add_contact_cables(h, m_contactCables.data(), (int32_t)m_contactCables.size());
    if (st == TGB_OK) st = tgb_finalize(h);
    if (st != TGB_OK) { tgb_destroy(h); check(st); }
    m_handle = h;
    m_cacheValid = false;
}
tgb_world* tgWorldB200Impl::handle() {
    finalize();
    return m_handle;
}
void tgWorldB200Impl::step(double dt) { run(dt, 1); }
void tgWorldB200Impl::run(double dt, int steps) {
    check(tgb_step(handle(), dt, steps));
    m_cacheValid = false;
}
void tgWorldB200Impl::stepWorld(double dt) {
    check(tgb_step_world(handle(), dt));
    m_cacheValid = false;
}
void tgWorldB200Impl::stepModels(double dt) {
    check(tgb_step_models(handle(), dt));
    m_cacheValid = false;
}
void tgWorldB200Impl::setMotorTorque(int cable, double torque) {
    if (m_torque.size() != m_cables.size()) m_torque.assign(m_cables.size(), 0.0);
    if (m_torque.at(cable) == torque && m_handle) return;
    m_torque[cable] = torque;
    check(tgb_set_controller(handle(), TGB_CTRL_TORQUE, m_torque.data(), (int32_t)m_torque.size(), 0, m_nEnv, 1));
}
void tgWorldB200Impl::refreshCache() {
    if (m_cacheValid) return;
    m_rodCache.resize((std::size_t)m_nEnv * m_bodies.size() * TGB_ROD_STATE_DOUBLES);
    m_cableCache.resize((std::size_t)m_nEnv * m_cables.size() * TGB_CABLE_OBS_DOUBLES);
    check(tgb_get_rod_state(handle(), 0, m_nEnv, m_rodCache.data()));
    if (!m_cables.empty()) check(tgb_get_cable_obs(handle(), 0, m_nEnv, m_cableCache.data()));
    m_ccCache.resize((std::size_t)m_nEnv * m_contactCables.size() * TGB_CCABLE_OBS_DOUBLES);
    if (!m_contactCables.empty()) check(tgb_get_contact_cable_obs(handle(), 0, m_nEnv, m_ccCache.data()));
    m_cacheValid = true;
}
static void anchorsOf(const std::vector<tgb_body_desc>& bodies, const tgb_cable_desc& c, btVector3& p1, btVector3& p2) {
    const tgb_body_desc& A = bodies[c.body_a];
    const tgb_body_desc& B = bodies[c.body_b];
    btTransform ta(btQuaternion(A.quat[0], A.quat[1], A.quat[2], A.quat[3]), btVector3(A.pos[0], A.pos[1], A.pos[2]));
    btTransform tb(btQuaternion(B.quat[0], B.quat[1], B.quat[2], B.quat[3]), btVector3(B.pos[0], B.pos[1], B.pos[2]));
    p1 = ta * btVector3(c.local_a[0], c.local_a[1], c.local_a[2]);
    p2 = tb * btVector3(c.local_b[0], c.local_b[1], c.local_b[2]);
}
static void contactAnchorsOf(const std::vector<tgb_body_desc>& bodies, const tgb_contact_cable_desc& c, btVector3& p1, btVector3& p2) {
    tgb_cable_desc plain = tgb_cable_desc();
    plain.body_a = c.body_a; plain.body_b = c.body_b;
    for (int k = 0; k < 3; ++k) { plain.local_a[k] = c.local_a[k]; plain.local_b[k] = c.local_b[k]; }
    anchorsOf(bodies, plain, p1, p2);
}
int tgWorldB200Impl::contactCableAnchors(int cable, int env) {
    if (!m_handle) return 0;
    refreshCache();
    return (int)m_ccCache[((std::size_t)env * m_contactCables.size() + (cable - kContactCableBase)) * TGB_CCABLE_OBS_DOUBLES + 4];
}
double tgWorldB200Impl::cableRestLength(int cable, int env) {
    if (cable >= kContactCableBase) {
        if (!m_handle) return m_contactCables.at(cable - kContactCableBase).rest_length;
        refreshCache();
        return m_ccCache[((std::size_t)env * m_contactCables.size() + (cable - kContactCableBase)) * TGB_CCABLE_OBS_DOUBLES + 0];
    }
    if (!m_handle) return m_cables.at(cable).rest_length;
    refreshCache();
    return m_cableCache[((std::size_t)env * m_cables.size() + cable) * TGB_CABLE_OBS_DOUBLES + 0];
}
double tgWorldB200Impl::cableLength(int cable, int env) {
    if (cable >= kContactCableBase) {
        if (!m_handle) { /* no sliding anchor yet: the straight line between the two ends */
            btVector3 p1, p2;
            contactAnchorsOf(m_bodies, m_contactCables.at(cable - kContactCableBase), p1, p2);
            return (p2 - p1).length();
        }
        refreshCache();
        return m_ccCache[((std::size_t)env * m_contactCables.size() + (cable - kContactCableBase)) * TGB_CCABLE_OBS_DOUBLES + 1];
    }
    if (!m_handle) {
        btVector3 p1, p2;
        anchorsOf(m_bodies, m_cables.at(cable), p1, p2);
        return (p2 - p1).length();
    }
    refreshCache();
    return m_cableCache[((std::size_t)env * m_cables.size() + cable) * TGB_CABLE_OBS_DOUBLES + 1];
}
double tgWorldB200Impl::cableTension(int cable, int env) {
    if (cable >= kContactCableBase) {
        if (!m_handle) {
            const tgb_contact_cable_desc& c = m_contactCables.at(cable - kContactCableBase);
            double t = (cableLength(cable, env) - c.rest_length) * c.stiffness;
            return t < 0.0 ? 0.0 : t;
        }
        refreshCache();
        return m_ccCache[((std::size_t)env * m_contactCables.size() + (cable - kContactCableBase)) * TGB_CCABLE_OBS_DOUBLES + 2];
    }
    if (!m_handle) {
        const tgb_cable_desc& c = m_cables.at(cable);
        double t = (cableLength(cable, env) - c.rest_length) * c.stiffness;
        return t < 0.0 ? 0.0 : t;
    }
    refreshCache();
    return m_cableCache[((std::size_t)env * m_cables.size() + cable) * TGB_CABLE_OBS_DOUBLES + 2];
}
double tgWorldB200Impl::cableMotorVelocity(int cable, int env) {
    if (cable >= kContactCableBase) return 0.0;
    if (!m_handle) return 0.0;
    refreshCache();
    return m_cableCache[((std::size_t)env * m_cables.size() + cable) * TGB_CABLE_OBS_DOUBLES + 3];
}
btVector3 tgWorldB200Impl::bodyPosition(int body, int env) {
    if (!m_handle) { const tgb_body_desc& b = m_bodies.at(body); return btVector3(b.pos[0], b.pos[1], b.pos[2]); }
    refreshCache();
    const double* s = &m_rodCache[((std::size_t)env * m_bodies.size() + body) * TGB_ROD_STATE_DOUBLES];
    return btVector3(s[0], s[1], s[2]);
}
btQuaternion tgWorldB200Impl::bodyOrientation(int body, int env) {
    if (!m_handle) { const tgb_body_desc& b = m_bodies.at(body); return btQuaternion(b.quat[0], b.quat[1], b.quat[2], b.quat[3]); }
    refreshCache();
    const double* s = &m_rodCache[((std::size_t)env * m_bodies.size() + body) * TGB_ROD_STATE_DOUBLES];
    return btQuaternion(s[3], s[4], s[5], s[6]);
}
void tgWorldB200Impl::setControlInput(int cable, double input, double dt) {
    if (input < 0.0) throw std::invalid_argument("Rest length is negative.");
    if (cable >= kContactCableBase) {
        const int i = cable - kContactCableBase;
        if (!m_handle) {
            /* before the device exists: moveMotors on the flat descriptor; no sliding anchor yet */
            tgb_contact_cable_desc& c = m_contactCables.at(i);
            double preferred = input;
            const double stepSize = c.target_velocity * dt, actual = cableLength(cable, 0);
            if ((actual - preferred) * c.stiffness > c.max_tension) preferred = actual - c.max_tension / c.stiffness;
            const double diff = preferred - c.rest_length;
            if (actual > c.min_actual_length || diff > 0) {
                if (std::fabs(diff) > stepSize) c.rest_length += (diff / std::fabs(diff)) * stepSize;
                else c.rest_length += diff;
            }
            c.rest_length = c.rest_length > c.min_rest_length ? c.rest_length : c.min_rest_length;
            return;
        }
        std::vector<double> in(m_contactCables.size(), std::nan(""));
        in.at(i) = input;
        check(tgb_set_contact_cable_control_input(m_handle, 0, m_nEnv, in.data(), 1, dt));
        m_cacheValid = false;
        return;
    }
    if (!m_handle) {
        /* before the device exists: tgBasicActuator::moveMotors on the flat descriptor (tgBasicActuator.cpp:170-225) */
        tgb_cable_desc& c = m_cables.at(cable);
        double preferred = input;
        double stepSize = c.target_velocity * dt;
        double actual = cableLength(cable, 0);
        if ((actual - preferred) * c.stiffness > c.max_tension) preferred = actual - c.max_tension / c.stiffness;
        double diff = preferred - c.rest_length;
        if (actual > c.min_actual_length || diff > 0) {
            if (std::fabs(diff) > stepSize) c.rest_length += (diff / std::fabs(diff)) * stepSize;
            else c.rest_length += diff;
        }
        c.rest_length = c.rest_length > c.min_rest_length ? c.rest_length : c.min_rest_length;
        return;
    }
    /* one cable, all environments: read-modify-write the per-cable input vector */
    std::vector<double> in(m_cables.size(), std::nan("")); /* NaN = leave this cable alone */
    in[cable] = input;
    check(tgb_set_control_input(m_handle, 0, m_nEnv, in.data(), 1, dt));
    m_cacheValid = false;
}

/* ---- tgWorld -------------------------------------------------------------------------------------- */
tgWorld::Config::Config(double g, double ws) : gravity(g), worldSize(ws) {
    if (ws <= 0.0) throw std::invalid_argument("worldSize is not positive");
}
tgWorld::tgWorld() : m_config(), m_pGround(new tgBoxGround()), m_pImpl(0), m_nEnv(1), m_precision(TGB_F32), m_device(0) {
    m_pImpl = new tgWorldB200Impl(m_config.gravity, m_pGround);
}
tgWorld::tgWorld(const Config& config) : m_config(config), m_pGround(new tgBoxGround()), m_pImpl(0), m_nEnv(1), m_precision(TGB_F32), m_device(0) {
    m_pImpl = new tgWorldB200Impl(m_config.gravity, m_pGround);
}
tgWorld::tgWorld(const Config& config, tgGround* ground)
    : m_config(config), m_pGround(ground), m_pImpl(0), m_nEnv(1), m_precision(TGB_F32), m_device(0) {
    m_pImpl = new tgWorldB200Impl(m_config.gravity, m_pGround);
}
tgWorld::~tgWorld() {
    delete m_pImpl;
    delete m_pGround;
}
void tgWorld::reset() {
    delete m_pImpl;
    m_pImpl = new tgWorldB200Impl(m_config.gravity, m_pGround);
    m_pImpl->setBatch(m_nEnv, m_precision, m_device);
}
void tgWorld::reset(const Config& config) { m_config = config; reset(); }
void tgWorld::reset(tgGround* ground) {
    delete m_pGround;
    m_pGround = ground;
    reset();
}
void tgWorld::step(double dt) const {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    m_pImpl->step(dt);
}
void tgWorld::setBatch(int n_env, tgb_precision precision, int device) {
    m_nEnv = n_env; m_precision = precision; m_device = device;
    m_pImpl->setBatch(n_env, precision, device);
}

/* ---- tgModel -------------------------------------------------------------------------------------- */
tgModel::tgModel() {}
tgModel::tgModel(const tgTags& tags) : tgTaggable(tags) {}
tgModel::~tgModel() {
    for (std::size_t i = 0; i < m_children.size(); ++i) delete m_children[i];
}
void tgModel::setup(tgWorld& world) {
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->setup(world);
}
void tgModel::teardown() {
    for (std::size_t i = 0; i < m_children.size(); ++i) {
        m_children[i]->teardown();
        delete m_children[i];
    }
    m_children.clear();
}
void tgModel::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->step(dt);
}
void tgModel::onVisit(const tgModelVisitor& r) const {
    r.render(*this);
    for (std::size_t i = 0; i < m_children.size(); ++i) m_children[i]->onVisit(r);
}
void tgModel::addChild(tgModel* pChild) {
    if (pChild == 0) throw std::invalid_argument("child is NULL");
    if (std::find(m_children.begin(), m_children.end(), pChild) != m_children.end()) throw std::invalid_argument("child is already a child");
    m_children.push_back(pChild);
}
std::string tgModel::toString(std::string prefix) const {
    std::ostringstream os;
    os << prefix << "tgModel(" << getTags() << ", " << m_children.size() << " children)";
    for (std::size_t i = 0; i < m_children.size(); ++i) os << "\n" << m_children[i]->toString(prefix + "  ");
    return os.str();
}
std::vector<tgSenseable*> tgModel::getSenseableDescendants() const {
    const std::vector<tgModel*> d = getDescendants();
    return std::vector<tgSenseable*>(d.begin(), d.end());
}
std::vector<tgModel*> tgModel::getDescendants() const {
    std::vector<tgModel*> out;
    for (std::size_t i = 0; i < m_children.size(); ++i) {
        out.push_back(m_children[i]);
        std::vector<tgModel*> sub = m_children[i]->getDescendants();
        out.insert(out.end(), sub.begin(), sub.end());
    }
    return out;
}

/* ---- rigids --------------------------------------------------------------------------------------- */
tgBaseRigid::tgBaseRigid(tgWorldB200Impl* world, int body, double mass, const tgTags& tags)
    : tgModel(tags), m_world(world), m_body(body), m_mass(mass) {
    if (!world || body < 0) throw std::invalid_argument("tgBaseRigid: no body");
}
btVector3 tgBaseRigid::centerOfMass(int env) const { return m_world->bodyPosition(m_body, env); }
btQuaternion tgBaseRigid::orientationQuaternion(int env) const { return m_world->bodyOrientation(m_body, env); }
btVector3 tgBaseRigid::orientation(int env) const {
    btMatrix3x3 rot(m_world->bodyOrientation(m_body, env));
    btScalar yaw = 0, pitch = 0, roll = 0;
    rot.getEulerYPR(yaw, pitch, roll);
    return btVector3(yaw, pitch, roll);
}
tgRod::Config::Config(double r, double d, double f, double rf, double res) : radius(r), density(d), friction(f), rollFriction(rf), restitution(res) {
    if (density < 0.0) throw std::range_error("Negative density");
    if (radius < 0.0) throw std::range_error("Negative radius");
    if (friction < 0.0) throw std::range_error("Negative friction");
    if (rollFriction < 0.0) throw std::range_error("Negative roll friction");
    if (restitution < 0.0 || restitution > 1.0) throw std::range_error("Restitution outside [0,1]");
}
tgRod::tgRod(tgWorldB200Impl* world, int body, double mass, const tgTags& tags, double length)
    : tgBaseRigid(world, body, mass, tags), m_length(length) {}

tgBox::Config::Config(double w, double h, double d, double f, double rf, double res)
    : width(w), height(h), density(d), friction(f), rollFriction(rf), restitution(res) {
    if (density < 0.0) throw std::range_error("Negative density");
    if (width < 0.0) throw std::range_error("Negative width");
    if (height < 0.0) throw std::range_error("Negative height");
    if (friction < 0.0) throw std::range_error("Negative friction");
    if (rollFriction < 0.0) throw std::range_error("Negative roll friction");
    if (restitution < 0.0) throw std::range_error("Negative restitution");
    if (restitution > 1.0) throw std::range_error("Restitution > 1");
}
tgBox::tgBox(tgWorldB200Impl* world, int body, double mass, const tgTags& tags, double length)
    : tgBaseRigid(world, body, mass, tags), m_length(length) {}

tgSphere::Config::Config(double r, double d, double f, double rf, double res) : radius(r), density(d), friction(f), rollFriction(rf), restitution(res) {
    if (density < 0.0) throw std::range_error("Negative density");
    if (radius < 0.0) throw std::range_error("Negative radius");
    if (friction < 0.0) throw std::range_error("Negative friction");
    if (rollFriction < 0.0) throw std::range_error("Negative roll friction");
    if (restitution < 0.0) throw std::range_error("Negative restitution");
    if (restitution > 1.0) throw std::range_error("Restitution > 1");
}
tgSphere::tgSphere(tgWorldB200Impl* world, int body, double mass, const tgTags& tags) : tgBaseRigid(world, body, mass, tags) {}

/* ---- actuators ------------------------------------------------------------------------------------ */
tgSpringCableActuator::Config::Config(double s, double d, double p, bool h, double mf, double tVel, double mnAL, double mnRL, double rot,
                                      bool moveCPA, bool moveCPB)
    : stiffness(s), damping(d), pretension(p), hist(h), maxTens(mf), targetVelocity(tVel), minActualLength(mnAL), minRestLength(mnRL),
      rotation(rot), moveCablePointAToEdge(moveCPA), moveCablePointBToEdge(moveCPB) {
    /* same checks and messages as tgSpringCableActuator.cpp:61-87 */
    if (s < 0.0) throw std::invalid_argument("stiffness is negative.");
    else if (d < 0.0) throw std::invalid_argument("damping is negative.");
    else if (mf < 0.0) throw std::invalid_argument("max tension is negative.");
    else if (mnAL < 0.0) throw std::invalid_argument("min Actual Length is negative.");
    else if (mnRL < 0.0) throw std::invalid_argument("min Rest Length is negative.");
    else if (std::fabs(rot) > M_PI * 2.0)
        throw std::invalid_argument("Abs of rotation is greater than 2pi. Are you sure you're setting the right parameters?");
}
void tgSpringCableActuator::Config::scale(double sf) {
    pretension *= sf; maxTens *= sf; targetVelocity *= sf; minActualLength *= sf; minRestLength *= sf;
}
tgSpringCableActuator::tgSpringCableActuator(tgWorldB200Impl* world, int cable, const tgTags& tags, const Config& config, double startLength)
    : tgModel(tags), m_world(world), m_cable(cable), m_config(config), m_startLength(startLength) {
    if (!world || cable < 0) throw std::invalid_argument("Pointer to tgBulletSpringCable is NULL.");
    if (m_config.targetVelocity < 0.0) throw std::invalid_argument("Target velocity is negative.");
    if (m_config.minActualLength < 0.0) throw std::invalid_argument("Minimum length is negative.");
}
tgSpringCableActuator::~tgSpringCableActuator() {}
double tgSpringCableActuator::getCurrentLength(int env) const { return m_world->cableLength(m_cable, env); }
double tgSpringCableActuator::getTension(int env) const { return m_world->cableTension(m_cable, env); }
double tgSpringCableActuator::getRestLength(int env) const { return m_world->cableRestLength(m_cable, env); }

tgBasicActuator::tgBasicActuator(tgWorldB200Impl* world, int cable, const tgTags& tags, const Config& config, double startLength)
    : tgSpringCableActuator(world, cable, tags, config, startLength), m_preferredLength(world->cableRestLength(cable, 0)) {}
void tgBasicActuator::setup(tgWorld& world) {
    notifySetup();
    tgModel::setup(world);
}
void tgBasicActuator::teardown() { tgModel::teardown(); }
void tgBasicActuator::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive.");
    notifyStep(dt); /* host observers; the cable force itself is part of the device step */
    tgModel::step(dt);
}
void tgBasicActuator::setControlInput(double input) {
    if (input < 0.0) throw std::invalid_argument("Rest length is negative.");
    m_preferredLength = input;
}
void tgBasicActuator::setControlInput(double input, double dt) {
    if (input < 0.0) throw std::invalid_argument("Rest length is negative.");
    m_preferredLength = input;
    moveMotors(dt);
}
void tgBasicActuator::moveMotors(double dt) { m_world->setControlInput(m_cable, m_preferredLength, dt); }

tgKinematicActuator::Config::Config(double s, double d, double p, double rad, double moFric, double moInert, bool back, bool h, double mf,
                                    double tVel, double mnAL, double mnRL, double rot)
    : tgSpringCableActuator::Config(s, d, p, h, mf, tVel, mnAL, mnRL, rot), radius(rad), motorFriction(moFric), motorInertia(moInert),
      backdrivable(back), maxOmega(tVel / rad), maxTorque(mf / rad) {
    if (rad <= 0.0) throw std::invalid_argument("Radius is non-positive");
    else if (moFric < 0.0) throw std::invalid_argument("Motor friction is negative.");
    else if (moInert <= 0.0) throw std::invalid_argument("Motor inertia is non-positive");
}
tgKinematicActuator::tgKinematicActuator(tgWorldB200Impl* world, int cable, const tgTags& tags, const Config& config, double startLength)
    : tgSpringCableActuator(world, cable, tags, config, startLength), m_kconfig(config), m_desiredTorque(0.0) {}
void tgKinematicActuator::setup(tgWorld& world) {
    notifySetup();
    tgModel::setup(world);
}
void tgKinematicActuator::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive.");
    notifyStep(dt);
    /* the desired torque of this step reaches the device before the models half runs; like the reference it lasts one step */
    m_world->setMotorTorque(m_cable, m_desiredTorque);
    tgModel::step(dt);
    m_desiredTorque = 0.0;
}
void tgKinematicActuator::setControlInput(double input) { m_desiredTorque = input; }
const double tgKinematicActuator::getVelocity() const { return m_world->cableMotorVelocity(m_cable, 0) * m_kconfig.radius; }

/* ---- view + simulation ---------------------------------------------------------------------------- */
tgSimView::tgSimView(tgWorld& world, double stepSize, double renderRate)
    : m_world(world), m_pSimulation(0), m_stepSize(stepSize), m_renderRate(renderRate), m_initialized(false) {
    if (stepSize < 0.0) throw std::invalid_argument("stepSize is not positive");
    if (renderRate < stepSize) throw std::invalid_argument("renderRate is less than stepSize");
}
tgSimView::~tgSimView() {
    if (m_pSimulation) teardown();
}
void tgSimView::bindToSimulation(tgSimulation& simulation) {
    if (m_pSimulation) throw std::invalid_argument("The view already belongs to a simulation.");
    m_pSimulation = &simulation;
}
void tgSimView::releaseFromSimulation() { m_pSimulation = 0; }
void tgSimView::setup() { m_initialized = true; }
void tgSimView::teardown() { m_initialized = false; }
void tgSimView::run() { run(10); }
void tgSimView::run(int steps) {
    if (!m_pSimulation) throw std::runtime_error("The view is not bound to a simulation");
    if (!m_pSimulation->hasHostObservers() && !m_pSimulation->hasHostStepCode()) {
        m_pSimulation->runFused(m_stepSize, steps);
        return;
    }
    for (int i = 0; i < steps; ++i) m_pSimulation->step(m_stepSize);
}
void tgSimView::reset() {
    if (m_pSimulation) m_pSimulation->reset();
}
void tgSimView::setStepSize(double stepSize) {
    if (stepSize <= 0.0) throw std::invalid_argument("stepSize is not positive");
    m_stepSize = stepSize;
}
void tgSimView::setRenderRate(double renderRate) { m_renderRate = renderRate < m_stepSize ? m_stepSize : renderRate; }

tgSimulation::tgSimulation(tgSimView& view) : m_view(view) {
    m_view.bindToSimulation(*this);
    m_view.setup();
}
tgSimulation::~tgSimulation() {
    teardown();
    m_view.releaseFromSimulation();
    for (std::size_t i = 0; i < m_models.size(); ++i) delete m_models[i];
    for (std::size_t i = 0; i < m_dataManagers.size(); ++i) delete m_dataManagers[i];
}
void tgSimulation::addDataManager(tgDataManager* pDataManager) {
    if (pDataManager == 0) throw std::invalid_argument("NULL pointer to tgDataManager");
    pDataManager->setup();
    m_dataManagers.push_back(pDataManager);
}
void tgSimulation::addModel(tgModel* pModel) {
    if (pModel == 0) throw std::invalid_argument("NULL pointer to tgModel");
    pModel->setup(m_view.world());
    m_models.push_back(pModel);
}
tgWorld& tgSimulation::getWorld() const { return m_view.world(); }
namespace {
template <class A>
bool observed(tgModel* m) {
    A* a = dynamic_cast<A*>(m);
    return a && a->numObservers() > 0;
}
}
bool tgSimulation::hasHostObservers() const {
    for (std::size_t i = 0; i < m_models.size(); ++i) {
        /* model-level observers cannot be detected generically (each model is its own tgSubject<T>); models
           that use them must step through step(dt). Actuator-level observers are detected here. */
        std::vector<tgModel*> all = m_models[i]->getDescendants();
        for (std::size_t j = 0; j < all.size(); ++j)
            if (observed<tgBasicActuator>(all[j]) || observed<tgKinematicActuator>(all[j])) return true;
    }
    return false;
}
bool tgSimulation::hasHostStepCode() const {
    for (std::size_t i = 0; i < m_models.size(); ++i)
        if (m_models[i]->stepsOnHost()) return true;
    return !m_dataManagers.empty(); /* loggers sample the state on the host after every step */
}
void tgSimulation::step(double dt) const {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    /* (A) world step of every environment on the device (tgSimulation.cpp:205) */
    m_view.world().b200().stepWorld(dt);
    /* (B) model tree walk on the host (tgSimulation.cpp:208-211): observers and the models' own step code see the state after
       the world step and may move motors through the C ABI, exactly where the reference runs them ... */
    for (std::size_t i = 0; i < m_models.size(); ++i) m_models[i]->step(dt);
    /* ... then the cable forces of every actuator's step on the device. A cable force depends only on body poses and its own
       rest length, so evaluating all of them after all observers equals the reference's per-actuator interleaving. */
    m_view.world().b200().stepModels(dt);
    /* data managers last (tgSimulation.cpp:213-216): loggers see the state after the complete step */
    for (std::size_t i = 0; i < m_dataManagers.size(); ++i) m_dataManagers[i]->step(dt);
}
void tgSimulation::runFused(double dt, int steps) const {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    m_view.world().b200().run(dt, steps);
}
void tgSimulation::run() const { m_view.run(); }
void tgSimulation::run(int steps) const { m_view.run(steps); }
void tgSimulation::teardown() {
    for (std::size_t i = 0; i < m_models.size(); ++i) m_models[i]->teardown();
    for (std::size_t i = 0; i < m_dataManagers.size(); ++i) m_dataManagers[i]->teardown();
}
void tgSimulation::reset() {
    teardown();
    m_view.world().reset();
    for (std::size_t i = 0; i < m_models.size(); ++i) m_models[i]->setup(m_view.world());
    /* after the models: the sensors are created from the rebuilt model tree (tgSimulation.cpp:147-156) */
    for (std::size_t i = 0; i < m_dataManagers.size(); ++i) m_dataManagers[i]->setup();
}
