// speck_b200/host/hostio_capi.cpp -- C ABI over the host loaders (include/speck_hostio.h).
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include "COO.h"
#include "CSR.h"
#include "Config.h"
#include "speck_hostio.h"

namespace {
thread_local std::string g_err;
int export_csr(const CSR<double> &m, size_t *rows, size_t *cols, size_t *nnz, uint32_t **rp, uint32_t **ci, double **v)
{
    *rows = m.rows; *cols = m.cols; *nnz = m.nnz;
    *rp = (uint32_t *)malloc((m.rows + 1) * sizeof(uint32_t));
    *ci = (uint32_t *)malloc((m.nnz ? m.nnz : 1) * sizeof(uint32_t));
    *v = (double *)malloc((m.nnz ? m.nnz : 1) * sizeof(double));
    memcpy(*rp, m.row_offsets.get(), (m.rows + 1) * sizeof(uint32_t));
    memcpy(*ci, m.col_ids.get(), m.nnz * sizeof(uint32_t));
    memcpy(*v, m.data.get(), m.nnz * sizeof(double));
    return 0;
}
bool key_of(const char *name, Config::Key &k)
{
    std::string s(name);
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    const struct { const char *n; Config::Key k; } table[] = {
        {"inputfile", Config::InputFile}, {"iterationswarmup", Config::IterationsWarmUp},
        {"iterationsexecution", Config::IterationsExecution}, {"trackindividualtimes", Config::TrackIndividualTimes},
        {"trackcompletetimes", Config::TrackCompleteTimes}, {"compareresult", Config::CompareResult},
        {"device", Config::Device}, {"devices", Config::Devices}, {"gpuconvert", Config::GpuConvert}};
    for (auto &e : table)
        if (s == e.n) { k = e.k; return true; }
    return false;
}
void init_cfg(const char *ini) { if (ini) Config::init(std::string(ini)); else Config::init(); }
}  // namespace

extern "C" {
const char *speck_host_last_error(void) { return g_err.c_str(); }

int speck_host_load_mtx_f64(const char *path, size_t *rows, size_t *cols, size_t *nnz, uint32_t **rp, uint32_t **ci, double **v)
{
    try {
        COO<double> coo = loadMTX<double>(path);
        CSR<double> csr;
        convert(csr, coo);
        return export_csr(csr, rows, cols, nnz, rp, ci, v);
    } catch (std::exception &e) { g_err = e.what(); return -1; }
}
int speck_host_load_hicsr_f64(const char *path, size_t *rows, size_t *cols, size_t *nnz, uint32_t **rp, uint32_t **ci, double **v)
{
    try {
        CSR<double> csr = loadCSR<double>(path);
        return export_csr(csr, rows, cols, nnz, rp, ci, v);
    } catch (std::exception &e) { g_err = e.what(); return -1; }
}
int speck_host_store_hicsr_f64(const char *path, size_t rows, size_t cols, size_t nnz, const uint32_t *rp, const uint32_t *ci, const double *v)
{
    try {
        CSR<double> m;
        m.alloc(rows, cols, nnz);
        memcpy(m.row_offsets.get(), rp, (rows + 1) * sizeof(uint32_t));
        memcpy(m.col_ids.get(), ci, nnz * sizeof(uint32_t));
        memcpy(m.data.get(), v, nnz * sizeof(double));
        storeCSR(m, path);
        return 0;
    } catch (std::exception &e) { g_err = e.what(); return -1; }
}
void speck_host_free(void *p) { free(p); }

int speck_host_config_get_int(const char *ini, const char *key, int fallback)
{
    Config::Key k;
    if (!key_of(key, k)) return fallback;
    init_cfg(ini);
    return Config::getInt(k, fallback);
}
int speck_host_config_get_bool(const char *ini, const char *key, int fallback)
{
    Config::Key k;
    if (!key_of(key, k)) return fallback;
    init_cfg(ini);
    return Config::getBool(k, fallback != 0) ? 1 : 0;
}
int speck_host_config_get_string(const char *ini, const char *key, const char *fallback, char *out, size_t cap)
{
    Config::Key k;
    std::string s = fallback ? fallback : "";
    if (key_of(key, k)) { init_cfg(ini); s = Config::getString(k, s); }
    if (cap == 0) return -1;
    strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
    return 0;
}
}
