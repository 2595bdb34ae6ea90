/* Stands in for src/tgcreator/tgBasicActuatorInfo.h of the reference: a model source that includes it gets tgBasicActuatorInfo of the B200 host API. */
#include "../../tgb_creator.h"
