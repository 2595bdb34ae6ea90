/* Stands in for src/core/tgPlaneGround.h of the reference: a model source that includes it gets tgPlaneGround of the B200 host API. */
#include "../../tgb_creator.h"
