/* Stands in for src/core/tgControllable.h of the reference: a model source that includes it gets tgControllable of the B200 host API. */
#include "../../tgb_creator.h"
