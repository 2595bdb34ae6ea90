/* Stands in for src/tgcreator/tgConnectorInfo.h of the reference: a model source that includes it gets tgConnectorInfo of the B200 host API. */
#include "../../tgb_creator.h"
