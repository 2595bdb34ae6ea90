/* Stands in for src/core/tgWorldImpl.h of the reference: a model source that includes it gets tgWorldImpl of the B200 host API. */
#include "../../tgb_creator.h"
