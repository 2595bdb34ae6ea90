/* Stands in for src/tgcreator/tgBuildSpec.h of the reference: a model source that includes it gets tgBuildSpec of the B200 host API. */
#include "../../tgb_creator.h"
