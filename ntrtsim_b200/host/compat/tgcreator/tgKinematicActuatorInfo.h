/* Stands in for src/tgcreator/tgKinematicActuatorInfo.h of the reference: a model source that includes it gets tgKinematicActuatorInfo of the B200 host API. */
#include "../../tgb_creator.h"
