/* Stands in for src/core/tgSimulation.h of the reference: a model source that includes it gets tgSimulation of the B200 host API. */
#include "../../tgb_creator.h"
