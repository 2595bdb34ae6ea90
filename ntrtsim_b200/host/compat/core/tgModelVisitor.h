/* Stands in for src/core/tgModelVisitor.h of the reference: a model source that includes it gets tgModelVisitor of the B200 host API. */
#include "../../tgb_creator.h"
