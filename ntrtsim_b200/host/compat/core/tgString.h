/* Stands in for src/core/tgString.h of the reference: a model source that includes it gets tgString (in tgb_creator.h) of the B200 host API. */
#include "../../tgb_creator.h"
