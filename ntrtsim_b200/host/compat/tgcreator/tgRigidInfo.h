/* Stands in for src/tgcreator/tgRigidInfo.h of the reference: a model source that includes it gets tgRigidInfo of the B200 host API. */
#include "../../tgb_creator.h"
