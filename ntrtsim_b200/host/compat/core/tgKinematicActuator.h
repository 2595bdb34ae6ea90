/* Stands in for src/core/tgKinematicActuator.h of the reference: a model source that includes it gets tgKinematicActuator of the B200 host API. */
#include "../../tgb_creator.h"
