/* Stands in for src/core/tgTags.h of the reference: a model source that includes it gets tgTags of the B200 host API. */
#include "../../tgb_creator.h"
