/* Stands in for src/core/tgCast.h of the reference: a model source that includes it gets tgCast of the B200 host API. */
#include "../../tgb_creator.h"
