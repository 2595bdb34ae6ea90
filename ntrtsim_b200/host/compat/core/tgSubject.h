/* Stands in for src/core/tgSubject.h of the reference: a model source that includes it gets tgSubject<T> of the B200 host API. */
#include "../../tgb_creator.h"
