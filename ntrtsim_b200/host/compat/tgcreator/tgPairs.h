/* Stands in for src/tgcreator/tgPairs.h of the reference: a model source that includes it gets tgPairs of the B200 host API. */
#include "../../tgb_creator.h"
