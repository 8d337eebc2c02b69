/*
 * rpt_errors.h — status codes shared by the host producers (rpt_host.h) and the CUDA
 * tracing backend (rpt_b200.h).
 *
 * The reference signals failure by silently returning (`World::from_path` -> None,
 * src/trace.rs:141-143), by panicking (`expect`, src/trace.rs:36) or by discarding the
 * result (`let _ =`, src/trace.rs:198,219-221).  None of those may cross an FFI boundary, so
 * every entry point returns one of these codes and the message is kept per context
 * (`rpt_last_error`).
 */
#ifndef RPT_ERRORS_H
#define RPT_ERRORS_H

enum {
    RPT_OK = 0,
    RPT_ERR_INVALID_ARGUMENT = -1, /* null pointer, zero size, index out of range */
    RPT_ERR_NO_DEVICE = -2,        /* no CUDA device / driver: there is no CPU fallback */
    RPT_ERR_CUDA = -3,             /* a CUDA runtime call or kernel failed (see rpt_last_error) */
    RPT_ERR_NOT_READY = -4,        /* world / config / rng not uploaded yet */
    RPT_ERR_RNG_DIMENSIONS = -5,   /* config needs more than the 31 usable R-sequence dimensions */
    RPT_ERR_SIZE_MISMATCH = -6,    /* buffer length != width*height of the current config */
    RPT_ERR_NCCL = -7,             /* NCCL unavailable or a collective failed */
    RPT_ERR_UNSUPPORTED = -8       /* e.g. BVH deeper than the traversal stack */
};

#endif /* RPT_ERRORS_H */
