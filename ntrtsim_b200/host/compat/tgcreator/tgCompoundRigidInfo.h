/* Stands in for src/tgcreator/tgCompoundRigidInfo.h of the reference: a model source that includes it gets tgCompoundRigidInfo of the B200 host API. */
#include "../../tgb_creator.h"
