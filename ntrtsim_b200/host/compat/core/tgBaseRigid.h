/* Stands in for src/core/tgBaseRigid.h of the reference: a model source that includes it gets tgBaseRigid of the B200 host API. */
#include "../../tgb_creator.h"
