#ifndef TGB_HOST_BT_SCALAR_H
#define TGB_HOST_BT_SCALAR_H
#include "btVector3.h"
#endif
