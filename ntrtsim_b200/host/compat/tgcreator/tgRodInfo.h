/* Stands in for src/tgcreator/tgRodInfo.h of the reference: a model source that includes it gets tgRodInfo of the B200 host API. */
#include "../../tgb_creator.h"
