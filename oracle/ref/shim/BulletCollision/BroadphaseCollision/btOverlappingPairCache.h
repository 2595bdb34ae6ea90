/* Shim (oracle/ref): see ref_collision_shim.h */
#include "ref_collision_shim.h"
