/* Stands in for src/core/tgRod.h of the reference: a model source that includes it gets tgRod of the B200 host API. */
#include "../../tgb_creator.h"
