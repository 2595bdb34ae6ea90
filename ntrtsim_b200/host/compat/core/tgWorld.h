/* Stands in for src/core/tgWorld.h of the reference: a model source that includes it gets tgWorld of the B200 host API. */
#include "../../tgb_creator.h"
