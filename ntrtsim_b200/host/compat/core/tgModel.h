/* Stands in for src/core/tgModel.h of the reference: a model source that includes it gets tgModel, tgModelVisitor hooks of the B200 host API. */
#include "../../tgb_creator.h"
