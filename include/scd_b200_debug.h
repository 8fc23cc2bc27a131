/*
 * scd_b200_debug.h -- entry points that exist ONLY in the debug build of the library
 * (python -m diffusion_models_dev_project_b200.build --debug -> _lib/libscd_b200_dbg.so, compiled with
 * -DSCD_DEBUG_STAMPS).  The default libscd_b200.so exports none of them and its kernels carry no stamp code.
 */
#ifndef SCD_B200_DEBUG_H
#define SCD_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif

/* Profiling aid (tools/timeline.py): when set to a device buffer of 8 x int64 per CTA, fp_march and
 * bp_tile record %globaltimer at their phase boundaries.  NULL (default) disables it.     */
void scd_debug_set_stamps(void *device_buffer);

#ifdef __cplusplus
}
#endif
#endif
