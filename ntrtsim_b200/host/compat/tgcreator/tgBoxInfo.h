/* Stands in for src/tgcreator/tgBoxInfo.h of the reference: a model source that includes it gets tgBoxInfo of the B200 host API. */
#include "../../tgb_creator.h"
