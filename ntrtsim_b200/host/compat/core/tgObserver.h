/* Stands in for src/core/tgObserver.h of the reference: a model source that includes it gets tgObserver<T> of the B200 host API. */
#include "../../tgb_creator.h"
