/* Stands in for src/core/tgBasicActuator.h of the reference: a model source that includes it gets tgBasicActuator of the B200 host API. */
#include "../../tgb_creator.h"
