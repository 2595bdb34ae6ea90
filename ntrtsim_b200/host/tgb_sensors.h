/*
 * tgb_sensors.h -- the "sensors" half of the host API: tgSensor / tgSensorInfo / tgDataManager / tgDataLogger2 with the reference's
 * names, call sequence and on-disk CSV schema (src/sensors/tgSensor.h, tgSensorInfo.h, tgRodSensor.cpp:73-186,
 * tgSpringCableActuatorSensor.cpp:60-150, tgDataManager.cpp:60-271, tgDataLogger2.cpp:46-226).  The sensors read the batched world
 * through the same getters as user code; `env` selects which environment a logger writes (one log file per logger, like the
 * reference's one process = one rollout).
 */
#ifndef TGB_SENSORS_H
#define TGB_SENSORS_H

#include "tgb_host.h"

#include <fstream>
#include <map>

/* One sensor on one senseable object (tgSensor.h:44-104): headings "type(tags).field", data as strings. */
class tgSensor {
public:
    explicit tgSensor(tgSenseable* pSens); /* NULL -> std::invalid_argument */
    virtual ~tgSensor() {}
    virtual std::vector<std::string> getSensorDataHeadings() = 0;
    virtual std::vector<std::string> getSensorData() = 0;
    void setEnv(int env) { m_env = env; }
protected:
    tgSenseable* m_pSens;
    int m_env;
};

/* Factory: decides whether a senseable gets sensors of its kind (tgSensorInfo.h:47-92). */
class tgSensorInfo {
public:
    virtual ~tgSensorInfo() {}
    virtual bool isThisMySenseable(tgSenseable* pSenseable) = 0;
    virtual std::vector<tgSensor*> createSensorsIfAppropriate(tgSenseable* pSenseable) = 0;
};

/* rod(tags).X Y Z Euler1 Euler2 Euler3 mass (tgRodSensor.cpp:73-186) */
class tgRodSensor : public tgSensor {
public:
    explicit tgRodSensor(tgRod* pRod);
    std::vector<std::string> getSensorDataHeadings();
    std::vector<std::string> getSensorData();
};
class tgRodSensorInfo : public tgSensorInfo {
public:
    bool isThisMySenseable(tgSenseable* pSenseable);
    std::vector<tgSensor*> createSensorsIfAppropriate(tgSenseable* pSenseable); /* not a tgRod -> std::invalid_argument */
};

/* SCA(tags).RestLen CurrLen Tension (tgSpringCableActuatorSensor.cpp:60-150) */
class tgSpringCableActuatorSensor : public tgSensor {
public:
    explicit tgSpringCableActuatorSensor(tgSpringCableActuator* pSCA);
    std::vector<std::string> getSensorDataHeadings();
    std::vector<std::string> getSensorData();
};
class tgSpringCableActuatorSensorInfo : public tgSensorInfo {
public:
    bool isThisMySenseable(tgSenseable* pSenseable);
    std::vector<tgSensor*> createSensorsIfAppropriate(tgSenseable* pSenseable);
};

/* compound(tag).X Y Z YawDegrees PitchDegrees RollDegrees mass for the rigids that share one "compound_XXXXXX" tag
   (tgCompoundRigidSensor.cpp:59-262): position = mean of the members' centres of mass, orientation = the first member's rotation
   since the sensor was made, mass = sum of the members' mass() -- each of which is the mass of the shared body, as in the reference. */
class tgCompoundRigidSensor : public tgSensor {
public:
    explicit tgCompoundRigidSensor(tgModel* pModel);               /* always std::invalid_argument: a tag is required */
    tgCompoundRigidSensor(tgModel* pModel, std::string tag);      /* no rigid with the tag -> std::runtime_error */
    std::vector<std::string> getSensorDataHeadings();
    std::vector<std::string> getSensorData();
    btVector3 getCenterOfMass();
    btVector3 getOrientation();
    double getMass();
private:
    std::string m_tag;
    std::vector<tgBaseRigid*> m_rigids;
    btQuaternion origOrientQuatInv;
};
/* One sensor per compound tag found among a model's descendants; tags already served are blacklisted so that the same compound is not
   sensed again through a child model (tgCompoundRigidSensorInfo.cpp:93-230). */
class tgCompoundRigidSensorInfo : public tgSensorInfo {
public:
    bool isThisMySenseable(tgSenseable* pSenseable);
    std::vector<tgSensor*> createSensorsIfAppropriate(tgSenseable* pSenseable);
private:
    std::map<std::string, int> getCompoundTags(tgModel* pModel);
    std::vector<std::string> blacklist;
};

/* Owns sensor infos and sensors; setup() creates the sensors of every senseable and its descendants (tgDataManager.cpp:86-139). */
class tgDataManager {
public:
    tgDataManager();
    virtual ~tgDataManager();
    virtual void setup();
    virtual void teardown();
    virtual void step(double dt); /* dt <= 0 -> std::invalid_argument */
    virtual void addSensorInfo(tgSensorInfo* pSensorInfo); /* takes ownership; NULL -> std::invalid_argument */
    virtual void addSenseable(tgSenseable* pSenseable);    /* NULL -> std::invalid_argument */
    virtual std::string toString() const;
    std::size_t numSensors() const { return m_sensors.size(); }
protected:
    void addSensorsHelper(tgSenseable* pSenseable);
    std::vector<tgSensor*> m_sensors;
    std::vector<tgSensorInfo*> m_sensorInfos;
    std::vector<tgSenseable*> m_senseables;
};

/* CSV logger (tgDataLogger2.cpp:46-226): file <prefix>_<mmddYYYY_HHMMSS>.txt, a first line of free text, a heading line
   "time,<i>_<heading>,...", then one line "time,v,v,...," per sample; a sample is taken when at least timeInterval has passed. */
class tgDataLogger2 : public tgDataManager {
public:
    tgDataLogger2(std::string fileNamePrefix, double timeInterval = 0.0, int env = 0);
    virtual ~tgDataLogger2();
    virtual void setup();
    virtual void teardown();
    virtual void step(double dt);
    virtual std::string toString() const;
    const std::string& getFileName() const { return m_fileName; }
protected:
    std::string m_fileNamePrefix, m_fileName;
    double m_timeInterval, m_totalTime, m_updateTime;
    int m_env;
    std::ofstream tgOutput;
};

#endif /* TGB_SENSORS_H */
