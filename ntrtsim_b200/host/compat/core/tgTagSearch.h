/* Stands in for src/core/tgTagSearch.h of the reference: a model source that includes it gets tgTagSearch of the B200 host API. */
#include "../../tgb_creator.h"
