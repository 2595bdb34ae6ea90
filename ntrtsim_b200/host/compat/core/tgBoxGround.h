/* Stands in for src/core/tgBoxGround.h of the reference: a model source that includes it gets tgBoxGround of the B200 host API. */
#include "../../tgb_creator.h"
