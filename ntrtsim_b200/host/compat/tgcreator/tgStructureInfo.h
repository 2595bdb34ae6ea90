/* Stands in for src/tgcreator/tgStructureInfo.h of the reference: a model source that includes it gets tgStructureInfo of the B200 host API. */
#include "../../tgb_creator.h"
