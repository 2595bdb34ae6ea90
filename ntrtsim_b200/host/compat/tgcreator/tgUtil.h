/* Stands in for src/tgcreator/tgUtil.h of the reference: a model source that includes it gets tgUtil of the B200 host API. */
#include "../../tgb_creator.h"
