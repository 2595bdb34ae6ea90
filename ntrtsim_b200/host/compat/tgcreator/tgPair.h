/* Stands in for src/tgcreator/tgPair.h of the reference: a model source that includes it gets tgPair of the B200 host API. */
#include "../../tgb_creator.h"
