/* Stands in for src/core/tgSphere.h of the reference: a model source that includes it gets tgSphere of the B200 host API. */
#include "../../tgb_creator.h"
