/*
 * ssnamg_io.h -- C ABI of libssnmat.so: host-side reader for the MAT-file level 5 inputs the
 * reference loads (`load('InputData/data1-...')`, Class1/APD_SsN_Class1.m:27;
 * `load('InputData/data4-...')`, Class2/APD_SsN_Class2.m:20).  SURVEY.md 8f row 4: the input path
 * of a standalone (non-MATLAB, non-Python) run.  Plain C, zlib is the only dependency; no CUDA.
 *
 * The bundled files are "MATLAB 5.0 MAT-file", little endian; every variable is a
 * zlib-compressed miMATRIX element whose numeric data MATLAB stored in the smallest integer type
 * that holds it (e.g. `p`, `q` as uint8, `m`, `n` as uint16, although their class is double).
 * ssn_mat_read_double() therefore widens whatever the stored type is to fp64, which is what
 * MATLAB's `load` hands the script.
 *
 * Variables the scripts use: Class 1 -- c (mn x 1), r (n x 1), l (m x 1), p (m x 1), q (n x 1),
 * gama (mn x 1), m, n; Class 2 has phi (mn x 1) and mu instead of gama (and the m x n matrix C).
 *
 * Every function returns SSN_MAT_OK (0) or a negative SSN_MAT_E_* code; nothing aborts.
 */
#ifndef SSNAMG_IO_H
#define SSNAMG_IO_H

#include <stdint.h>

#if defined(__GNUC__)
#define SSN_IO_API __attribute__((visibility("default")))
#else
#define SSN_IO_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SSN_MAT_OK            = 0,
    SSN_MAT_E_IO          = -1,  /* file missing or short read                                     */
    SSN_MAT_E_FORMAT      = -2,  /* not a little-endian level 5 MAT-file, or a truncated element    */
    SSN_MAT_E_ZLIB        = -3,  /* a compressed element does not inflate                           */
    SSN_MAT_E_NOMEM       = -4,
    SSN_MAT_E_INVALID     = -5,  /* null pointer / index out of range                               */
    SSN_MAT_E_UNSUPPORTED = -6   /* sparse, char, struct, cell, complex or N-D variable             */
};

typedef struct ssn_mat ssn_mat;

/* Reads and indexes the whole file (variables are inflated once, here). */
SSN_IO_API int ssn_mat_open(const char* path, ssn_mat** out);
SSN_IO_API void ssn_mat_close(ssn_mat* m);

/* Number of top-level variables, in file order; name of the i-th (NULL if out of range). */
SSN_IO_API int ssn_mat_count(const ssn_mat* m);
SSN_IO_API const char* ssn_mat_name(const ssn_mat* m, int i);
/* Index of the variable called `name`, or -1. */
SSN_IO_API int ssn_mat_find(const ssn_mat* m, const char* name);

/* Dimensions of the i-th variable (rows/cols may be NULL).  SSN_MAT_E_UNSUPPORTED when the
 * variable is not a real full numeric 2-D array; the dimensions are still filled in when known. */
SSN_IO_API int ssn_mat_dims(const ssn_mat* m, int i, int64_t* rows, int64_t* cols);

/* Copies the i-th variable into out[rows*cols], column-major, widened to fp64. */
SSN_IO_API int ssn_mat_read_double(const ssn_mat* m, int i, double* out);

/* One OT problem as the scripts see it after `load`: all pointers are malloc'ed by
 * ssn_problem_load and released by ssn_problem_free.  Optional variables that the file does not
 * hold are NULL (gama: the Class 2 file has none; phi: the Class 1 file has none) or NaN (mu).
 * gama may hold +Inf (it does in data1-500.mat). */
typedef struct {
    int64_t m, n;
    double* c;      /* mn, column-major m x n cost                                                  */
    double* r;      /* n  column marginals                                                          */
    double* l;      /* m  row marginals                                                             */
    double* p;      /* m                                                                            */
    double* q;      /* n                                                                            */
    double* gama;   /* mn upper bounds (Class 1) or NULL                                            */
    double* phi;    /* mn (Class 2) or NULL                                                         */
    double  mu;     /* transported mass (Class 2) or NaN                                            */
} ssn_problem;

/* load + size checks (the sizes the scripts assume: numel(c) = m*n, numel(l) = numel(p) = m,
 * numel(r) = numel(q) = n).  m and n are taken from the variables `m`, `n` when present, else
 * from numel(l), numel(r). */
SSN_IO_API int ssn_problem_load(const char* path, ssn_problem* out);
SSN_IO_API void ssn_problem_free(ssn_problem* pb);

/* Text for a status code (static storage). */
SSN_IO_API const char* ssn_mat_strerror(int status);

#ifdef __cplusplus
}
#endif
#endif
