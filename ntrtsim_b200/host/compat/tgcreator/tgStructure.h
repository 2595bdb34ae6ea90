/* Stands in for src/tgcreator/tgStructure.h of the reference: a model source that includes it gets tgStructure of the B200 host API. */
#include "../../tgb_creator.h"
