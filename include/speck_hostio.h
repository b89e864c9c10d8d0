/*
 * speck_hostio.h -- C ABI over the host-side loaders of the spECK driver, so that bindings and
 * the CPU test-suite can exercise them without a GPU.  Replaces (reference paths):
 *   speck_host_load_mtx_f64    <- loadMTX + convert(CSR, COO), source/COO.cpp:52-164, source/CSR.cpp:173-212
 *   speck_host_load_hicsr_f64  <- loadCSR,  source/CSR.cpp:88-120
 *   speck_host_store_hicsr_f64 <- storeCSR, source/CSR.cpp:122-137
 *   speck_host_config_*        <- Config::init/getInt/getBool/getString, source/Config.cpp:4-40
 * Arrays returned through ** are malloc'ed; release with speck_host_free.  0 = success.
 */
#ifndef SPECK_HOSTIO_H
#define SPECK_HOSTIO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
const char *speck_host_last_error(void);
int speck_host_load_mtx_f64(const char *path, size_t *rows, size_t *cols, size_t *nnz, uint32_t **row_offsets,
                            uint32_t **col_ids, double **data);
int speck_host_load_hicsr_f64(const char *path, size_t *rows, size_t *cols, size_t *nnz, uint32_t **row_offsets,
                              uint32_t **col_ids, double **data);
int speck_host_store_hicsr_f64(const char *path, size_t rows, size_t cols, size_t nnz, const uint32_t *row_offsets,
                               const uint32_t *col_ids, const double *data);
void speck_host_free(void *p);
/* ini_path may be NULL (defaults only).  key is the ini key name, case-insensitive. */
int speck_host_config_get_int(const char *ini_path, const char *key, int fallback);
int speck_host_config_get_bool(const char *ini_path, const char *key, int fallback);
int speck_host_config_get_string(const char *ini_path, const char *key, const char *fallback, char *out, size_t cap);
#ifdef __cplusplus
}
#endif
#endif
