/* Stands in for src/core/tgTaggable.h of the reference: a model source that includes it gets tgTaggable of the B200 host API. */
#include "../../tgb_creator.h"
