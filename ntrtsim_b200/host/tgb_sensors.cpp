/*
 * tgb_sensors.cpp -- sensors, data managers and the CSV logger of the host API; see tgb_sensors.h for the reference lines.
 */
#include "tgb_sensors.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace {
/* operator<<(ostream, tgTags): tags joined by single spaces (tgTags.h:315-326) */
std::string joinTags(const tgTags& tags) {
    std::ostringstream os;
    const std::deque<std::string>& t = tags.getTags();
    for (std::size_t i = 0; i < t.size(); ++i) {
        if (i != 0) os << " ";
        os << t[i];
    }
    return os.str();
}
/* the sensors print numbers through a default-formatted stringstream (6 significant digits) */
std::string number(double v) {
    std::stringstream ss;
    ss << v;
    return ss.str();
}
} // namespace

tgSensor::tgSensor(tgSenseable* pSens) : m_pSens(pSens), m_env(0) {
    if (pSens == 0) throw std::invalid_argument("Pointer to pSens is NULL inside tgSensor.");
}

/* ---- rods ------------------------------------------------------------------------------------------------- */
tgRodSensor::tgRodSensor(tgRod* pRod) : tgSensor(pRod) {
    if (pRod == 0) throw std::invalid_argument("Pointer to pRod is NULL inside tgRodSensor.");
}
std::vector<std::string> tgRodSensor::getSensorDataHeadings() {
    tgRod* rod = dynamic_cast<tgRod*>(m_pSens);
    const std::string prefix = "rod(" + joinTags(rod->getTags());
    const char* fields[7] = {").X", ").Y", ").Z", ").Euler1", ").Euler2", ").Euler3", ").mass"};
    std::vector<std::string> headings;
    for (int i = 0; i < 7; ++i) headings.push_back(prefix + fields[i]);
    return headings;
}
std::vector<std::string> tgRodSensor::getSensorData() {
    tgRod* rod = dynamic_cast<tgRod*>(m_pSens);
    const btVector3 com = rod->centerOfMass(m_env);
    const btVector3 orient = rod->orientation(m_env);
    std::vector<std::string> data;
    for (int i = 0; i < 3; ++i) data.push_back(number(com[i]));
    for (int i = 0; i < 3; ++i) data.push_back(number(orient[i]));
    data.push_back(number(rod->mass()));
    return data;
}
bool tgRodSensorInfo::isThisMySenseable(tgSenseable* pSenseable) { return dynamic_cast<tgRod*>(pSenseable) != 0; }
std::vector<tgSensor*> tgRodSensorInfo::createSensorsIfAppropriate(tgSenseable* pSenseable) {
    if (!isThisMySenseable(pSenseable)) throw std::invalid_argument("pSenseable is NOT a tgRod, inside tgRodSensorInfo.");
    return std::vector<tgSensor*>(1, new tgRodSensor(dynamic_cast<tgRod*>(pSenseable)));
}

/* ---- spring-cable actuators ------------------------------------------------------------------------------- */
tgSpringCableActuatorSensor::tgSpringCableActuatorSensor(tgSpringCableActuator* pSCA) : tgSensor(pSCA) {
    if (pSCA == 0) throw std::invalid_argument("Pointer to pSCA is NULL inside tgSpringCableActuatorSensor.");
}
std::vector<std::string> tgSpringCableActuatorSensor::getSensorDataHeadings() {
    tgSpringCableActuator* sca = dynamic_cast<tgSpringCableActuator*>(m_pSens);
    const std::string prefix = "SCA(" + joinTags(sca->getTags());
    std::vector<std::string> headings;
    headings.push_back(prefix + ").RestLen");
    headings.push_back(prefix + ").CurrLen");
    headings.push_back(prefix + ").Tension");
    return headings;
}
std::vector<std::string> tgSpringCableActuatorSensor::getSensorData() {
    tgSpringCableActuator* sca = dynamic_cast<tgSpringCableActuator*>(m_pSens);
    std::vector<std::string> data;
    data.push_back(number(sca->getRestLength(m_env)));
    data.push_back(number(sca->getCurrentLength(m_env)));
    data.push_back(number(sca->getTension(m_env)));
    return data;
}
bool tgSpringCableActuatorSensorInfo::isThisMySenseable(tgSenseable* pSenseable) { return dynamic_cast<tgSpringCableActuator*>(pSenseable) != 0; }
std::vector<tgSensor*> tgSpringCableActuatorSensorInfo::createSensorsIfAppropriate(tgSenseable* pSenseable) {
    if (!isThisMySenseable(pSenseable)) throw std::invalid_argument("pSenseable is NOT a tgSpringCableActuator, inside tgSpringCableActuatorSensorInfo.");
    return std::vector<tgSensor*>(1, new tgSpringCableActuatorSensor(dynamic_cast<tgSpringCableActuator*>(pSenseable)));
}

/* ---- compound rigid bodies --------------------------------------------------------------------------------- */
tgCompoundRigidSensor::tgCompoundRigidSensor(tgModel* pModel) : tgSensor(pModel) {
    throw std::invalid_argument("Called the constructor for tgCompoundRigidSensor with only a pointer, NOT a tag. That's not allowed, a specific compound tag must be passed in also.");
}
tgCompoundRigidSensor::tgCompoundRigidSensor(tgModel* pModel, std::string tag) : tgSensor(pModel), m_tag(tag) {
    if (pModel == 0) throw std::invalid_argument("Pointer to pModel is NULL inside tgCompoundRigidSensor.");
    m_rigids = pModel->find<tgBaseRigid>(m_tag);
    if (m_rigids.empty()) throw std::runtime_error("tgCompoundRigidSensor found no rigid bodies with its tag - something is wrong (inside constructor.)");
    origOrientQuatInv = m_rigids[0]->orientationQuaternion(0).inverse();
}
btVector3 tgCompoundRigidSensor::getCenterOfMass() {
    btVector3 com(0.0, 0.0, 0.0);
    for (std::size_t i = 0; i < m_rigids.size(); ++i) com += m_rigids[i]->centerOfMass(m_env);
    com /= (btScalar)m_rigids.size();
    return com;
}
btVector3 tgCompoundRigidSensor::getOrientation() {
    const btQuaternion diff = m_rigids[0]->orientationQuaternion(m_env) * origOrientQuatInv;
    btMatrix3x3 rot(diff);
    btScalar yaw = 0.0, pitch = 0.0, roll = 0.0;
    rot.getEulerYPR(yaw, pitch, roll);
    return btVector3(180 / M_PI * yaw, 180 / M_PI * pitch, 180 / M_PI * roll);
}
double tgCompoundRigidSensor::getMass() {
    double mass = 0;
    for (std::size_t i = 0; i < m_rigids.size(); ++i) mass += m_rigids[i]->mass();
    return mass;
}
std::vector<std::string> tgCompoundRigidSensor::getSensorDataHeadings() {
    const std::string prefix = "compound(" + m_tag;
    const char* fields[7] = {").X", ").Y", ").Z", ").YawDegrees", ").PitchDegrees", ").RollDegrees", ").mass"};
    std::vector<std::string> headings;
    for (int i = 0; i < 7; ++i) headings.push_back(prefix + fields[i]);
    return headings;
}
std::vector<std::string> tgCompoundRigidSensor::getSensorData() {
    const btVector3 com = getCenterOfMass(), orient = getOrientation();
    std::vector<std::string> data;
    for (int i = 0; i < 3; ++i) data.push_back(number(com[i]));
    for (int i = 0; i < 3; ++i) data.push_back(number(orient[i]));
    data.push_back(number(getMass()));
    return data;
}
namespace {
/* "compound_" followed by six word characters */
bool findCompoundTag(const std::string& tags, std::string& matched) {
    const std::string prefix = "compound_";
    for (std::size_t pos = tags.find(prefix); pos != std::string::npos; pos = tags.find(prefix, pos + prefix.size())) {
        if (tags.size() < pos + prefix.size() + 6) continue;
        bool ok = true;
        for (std::size_t i = 0; i < 6; ++i) {
            const unsigned char ch = (unsigned char)tags[pos + prefix.size() + i];
            ok = ok && (std::isalnum(ch) || ch == '_');
        }
        if (ok) { matched = tags.substr(pos, prefix.size() + 6); return true; }
    }
    return false;
}
} // namespace
std::map<std::string, int> tgCompoundRigidSensorInfo::getCompoundTags(tgModel* pModel) {
    std::map<std::string, int> compounds;
    std::vector<tgModel*> descendants = pModel->getDescendants();
    for (std::size_t i = 0; i < descendants.size(); ++i) {
        std::string matched;
        if (!findCompoundTag(joinTags(descendants[i]->getTags()), matched)) continue;
        if (dynamic_cast<tgBaseRigid*>(descendants[i]) == 0)
            throw std::runtime_error("A tgModel that is NOT a rigid body has a compound tag attached to it. Only rigid bodies should have compound tags.");
        compounds[matched] += 1;
    }
    return compounds;
}
bool tgCompoundRigidSensorInfo::isThisMySenseable(tgSenseable* pSenseable) {
    if (pSenseable == 0) throw std::invalid_argument("pSenseable was NULL inside tgCompoundRigidSensorInfo.");
    tgModel* pModel = dynamic_cast<tgModel*>(pSenseable);
    if (pModel == 0) return false;
    const std::map<std::string, int> compounds = getCompoundTags(pModel);
    for (std::map<std::string, int>::const_iterator it = compounds.begin(); it != compounds.end(); ++it) {
        if (it->second >= 2) {
            if (std::find(blacklist.begin(), blacklist.end(), it->first) == blacklist.end()) return true;
        } else if (it->second == 1) {
            throw std::runtime_error("There is only one object with a compound tag. That's not possible, compound tgModels should be at least two objects. Exiting.");
        }
    }
    return false;
}
std::vector<tgSensor*> tgCompoundRigidSensorInfo::createSensorsIfAppropriate(tgSenseable* pSenseable) {
    if (!isThisMySenseable(pSenseable))
        throw std::invalid_argument("pSenseable is NOT a valid tgModel or does not have non-blacklisted compound rigid bodies, inside tgCompoundRigidSensorInfo.");
    tgModel* pModel = dynamic_cast<tgModel*>(pSenseable);
    std::map<std::string, int> compounds = getCompoundTags(pModel);
    for (std::size_t i = 0; i < blacklist.size(); ++i) compounds.erase(blacklist[i]);
    std::vector<tgSensor*> sensors;
    for (std::map<std::string, int>::const_iterator it = compounds.begin(); it != compounds.end(); ++it) {
        sensors.push_back(new tgCompoundRigidSensor(pModel, it->first));
        blacklist.push_back(it->first);
    }
    return sensors;
}

/* ---- data manager ------------------------------------------------------------------------------------------ */
tgDataManager::tgDataManager() {}
tgDataManager::~tgDataManager() {
    for (std::size_t i = 0; i < m_sensors.size(); ++i) delete m_sensors[i];
    for (std::size_t i = 0; i < m_sensorInfos.size(); ++i) delete m_sensorInfos[i];
}
void tgDataManager::addSensorsHelper(tgSenseable* pSenseable) {
    for (std::size_t i = 0; i < m_sensorInfos.size(); ++i) {
        if (!m_sensorInfos[i]->isThisMySenseable(pSenseable)) continue;
        std::vector<tgSensor*> created = m_sensorInfos[i]->createSensorsIfAppropriate(pSenseable);
        for (std::size_t k = 0; k < created.size(); ++k)
            if (created[k] != 0) m_sensors.push_back(created[k]);
    }
}
void tgDataManager::setup() {
    for (std::size_t j = 0; j < m_senseables.size(); ++j) {
        addSensorsHelper(m_senseables[j]);
        std::vector<tgSenseable*> descendants = m_senseables[j]->getSenseableDescendants();
        for (std::size_t k = 0; k < descendants.size(); ++k) addSensorsHelper(descendants[k]);
    }
}
void tgDataManager::teardown() {
    for (std::size_t i = 0; i < m_sensors.size(); ++i) delete m_sensors[i];
    m_sensors.clear();
}
void tgDataManager::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
}
void tgDataManager::addSensorInfo(tgSensorInfo* pSensorInfo) {
    if (pSensorInfo == 0) throw std::invalid_argument("pSensorInfo is NULL inside tgDataManager::addSensorInfo");
    m_sensorInfos.push_back(pSensorInfo);
}
void tgDataManager::addSenseable(tgSenseable* pSenseable) {
    if (pSenseable == 0) throw std::invalid_argument("pSenseable is NULL inside tgDataManager::addSenseable");
    m_senseables.push_back(pSenseable);
}
std::string tgDataManager::toString() const {
    std::ostringstream os;
    os << "tgDataManager with " << m_sensors.size() << " sensors, " << m_sensorInfos.size() << " sensorInfos, and " << m_senseables.size()
       << " senseable objects." << std::endl;
    return os.str();
}

/* ---- CSV logger --------------------------------------------------------------------------------------------- */
tgDataLogger2::tgDataLogger2(std::string fileNamePrefix, double timeInterval, int env)
    : m_fileNamePrefix(fileNamePrefix), m_timeInterval(timeInterval), m_totalTime(0.0), m_updateTime(0.0), m_env(env) {
    if (m_fileNamePrefix == "")
        throw std::invalid_argument("File name cannot be the empty string. Please pass in a path to a file that can be opened.");
    if (m_timeInterval < 0.0) throw std::invalid_argument("Time interval must be nonnegative. Negative time intervals do not make sense.");
    if (env < 0) throw std::invalid_argument("Environment index must be nonnegative.");
    if (m_fileNamePrefix.at(0) == '~') { /* "~" expands to $HOME */
        const char* home = std::getenv("HOME");
        m_fileNamePrefix.erase(0, 1);
        m_fileNamePrefix = std::string(home ? home : "") + m_fileNamePrefix;
    }
}
tgDataLogger2::~tgDataLogger2() {}
void tgDataLogger2::setup() {
    tgDataManager::setup();
    for (std::size_t i = 0; i < m_sensors.size(); ++i) m_sensors[i]->setEnv(m_env);
    time_t rawtime;
    char fileTime[64];
    time(&rawtime);
    strftime(fileTime, sizeof(fileTime), "%m%d%Y_%H%M%S", localtime(&rawtime));
    m_fileName = m_fileNamePrefix + "_" + fileTime + ".txt";
    std::cout << "tgDataLogger2 will be saving data to the file: " << std::endl << m_fileName << std::endl;
    tgOutput.open(m_fileName.c_str());
    if (!tgOutput.is_open())
        throw std::runtime_error("Log file could not be opened. Usually, this is because the directory you specified does not exist. Check for spelling errors.");
    tgOutput << "tgDataLogger2 started logging at time " << fileTime << ", with " << m_sensors.size() << " sensors on " << m_senseables.size()
             << " senseable objects." << std::endl;
    tgOutput << "time,";
    for (std::size_t i = 0; i < m_sensors.size(); ++i) {
        std::vector<std::string> headings = m_sensors[i]->getSensorDataHeadings();
        for (std::size_t j = 0; j < headings.size(); ++j) tgOutput << i << "_" << headings[j] << ",";
    }
    tgOutput << std::endl;
    tgOutput.close();
    m_totalTime = 0.0;
    m_updateTime = 0.0;
}
void tgDataLogger2::teardown() {
    tgDataManager::teardown();
    tgOutput.close();
}
void tgDataLogger2::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    m_totalTime += dt;
    m_updateTime += dt;
    if (m_updateTime >= m_timeInterval) {
        tgOutput.open(m_fileName.c_str(), std::ios::app);
        tgOutput << m_totalTime << ",";
        for (std::size_t i = 0; i < m_sensors.size(); ++i) {
            std::vector<std::string> data = m_sensors[i]->getSensorData();
            for (std::size_t j = 0; j < data.size(); ++j) tgOutput << data[j] << ",";
        }
        tgOutput << std::endl;
        tgOutput.close();
        m_updateTime = 0.0;
    }
}
std::string tgDataLogger2::toString() const {
    std::ostringstream os;
    os << tgDataManager::toString() << "This tgDataManager is a tgDataLogger2. " << std::endl;
    return os.str();
}
