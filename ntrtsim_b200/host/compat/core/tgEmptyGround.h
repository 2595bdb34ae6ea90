/* Stands in for src/core/tgEmptyGround.h of the reference: a model source that includes it gets tgEmptyGround of the B200 host API. */
#include "../../tgb_creator.h"
