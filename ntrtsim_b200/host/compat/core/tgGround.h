/* Stands in for src/core/tgGround.h of the reference: a model source that includes it gets tgGround of the B200 host API. */
#include "../../tgb_creator.h"
