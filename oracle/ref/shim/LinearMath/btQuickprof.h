/* Shim: Bullet's scoped profiler is a no-op here (the reference wraps its uses in #ifndef BT_NO_PROFILE). */
#ifndef REF_SHIM_BT_QUICKPROF_H
#define REF_SHIM_BT_QUICKPROF_H
#ifndef BT_NO_PROFILE
#define BT_NO_PROFILE 1
#endif
#define BT_PROFILE(name)
#endif
