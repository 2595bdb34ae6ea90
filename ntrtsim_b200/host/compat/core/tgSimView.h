/* Stands in for src/core/tgSimView.h of the reference: a model source that includes it gets tgSimView of the B200 host API. */
#include "../../tgb_creator.h"
