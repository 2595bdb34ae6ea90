/* Stands in for src/tgcreator/tgNodes.h of the reference: a model source that includes it gets tgNodes of the B200 host API. */
#include "../../tgb_creator.h"
