/* Shim (oracle/ref): nothing of btDefaultMotionState is used by the sources compiled here */
#include "LinearMath/btTransform.h"
