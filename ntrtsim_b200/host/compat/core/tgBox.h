/* Stands in for src/core/tgBox.h of the reference: a model source that includes it gets tgBox of the B200 host API. */
#include "../../tgb_creator.h"
