/* Stands in for src/core/tgSenseable.h of the reference: a model source that includes it gets tgSenseable of the B200 host API. */
#include "../../tgb_creator.h"
