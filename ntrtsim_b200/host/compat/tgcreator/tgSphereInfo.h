/* Stands in for src/tgcreator/tgSphereInfo.h of the reference: a model source that includes it gets tgSphereInfo of the B200 host API. */
#include "../../tgb_creator.h"
