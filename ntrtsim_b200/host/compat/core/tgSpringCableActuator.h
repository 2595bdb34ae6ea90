/* Stands in for src/core/tgSpringCableActuator.h of the reference: a model source that includes it gets tgSpringCableActuator of the B200 host API. */
#include "../../tgb_creator.h"
