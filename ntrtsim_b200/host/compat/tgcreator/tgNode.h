/* Stands in for src/tgcreator/tgNode.h of the reference: a model source that includes it gets tgNode of the B200 host API. */
#include "../../tgb_creator.h"
