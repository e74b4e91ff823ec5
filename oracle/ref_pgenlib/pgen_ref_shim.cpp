// Test infrastructure (oracle/): a C shim over the REFERENCE's own vendored pgenlib, compiled from the sources where they
// lie under /root/reference/external_libs/pgenlib (recipe: oracle/build_native.py -> oracle/_ref/libpgenlib_ref.so).
// It calls the library exactly as the reference does - PgenReader::Load(file, n_samples, subset_1based, threads), then
// PgenReader::ReadHardcalls(buf, n, thread, variant, allele_idx = 1) per variant (src/Geno.cpp:1089, :1796-1798, :2572-2574)
// - so the .pgen restatement (oracle/pgen.py), the host decoder (host/pgen.cpp) and the device decoder (csrc/pgen_core.h)
// are pinned against reference code on files holding every record type, not only on the reference's one fixture.
// Only tests/ and bench.py's cpu_baseline leg load this; nothing under regenie_b200/ does.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "pgenlibr.h"

extern "C" {

// out[v][i] = what the reference's reader leaves in `g`: ALT allele count 0 / 1 / 2, missing = -3.  subset_1based may be
// NULL (all samples).  Returns 0, or -1 with a message in err.  seconds (optional) = time inside the Read loop only.
int pgref_read_hardcalls(const char* path, uint32_t n_raw, const int* subset_1based, uint32_t n_subset, uint32_t v0, uint32_t nv,
                         double* out, double* seconds, char* err, int err_len) {
  try {
    PgenReader pgr;
    std::vector<int> subset;
    if (subset_1based) subset.assign(subset_1based, subset_1based + n_subset);
    pgr.Load(path, n_raw, subset, 1);
    if (pgr.GetRawSampleCt() != n_raw) throw std::string("sample count mismatch");
    if (v0 + nv > pgr.GetVariantCt()) throw std::string("variant range out of bounds");
    if (pgr.GetMaxAlleleCt() != 2) throw std::string("multiallelic file");
    const size_t n = subset_1based ? n_subset : n_raw;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t v = 0; v < nv; ++v) pgr.ReadHardcalls(out + (size_t)v * n, n, 0, (int)(v0 + v), 1);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    pgr.Close();
    return 0;
  } catch (const std::string& e) {
    if (err && err_len > 0) { strncpy(err, e.c_str(), (size_t)err_len - 1); err[err_len - 1] = 0; }
    return -1;
  } catch (const char* e) {
    if (err && err_len > 0) { strncpy(err, e, (size_t)err_len - 1); err[err_len - 1] = 0; }
    return -1;
  } catch (...) {
    if (err && err_len > 0) { strncpy(err, "pgenlib raised an exception", (size_t)err_len - 1); err[err_len - 1] = 0; }
    return -1;
  }
}

// pgenlib's own structural validation of the whole file (PgrValidate: record types, difflist group byte counts, trailing
// bits): 0 = the file is well-formed by the reference library's rules.
int pgref_validate(const char* path, char* err, int err_len) {
  namespace p2 = plink2;
  p2::PgenFileInfo pgfi;
  p2::PreinitPgfi(&pgfi);
  p2::PgenHeaderCtrl header_ctrl;
  uintptr_t alloc_cacheline_ct = 0;
  char errbuf[p2::kPglErrstrBufBlen];
  errbuf[0] = 0;
  int rc = -1;
  unsigned char* pgfi_alloc = nullptr;
  unsigned char* pgr_alloc = nullptr;
  uintptr_t* genovec = nullptr;
  std::vector<uintptr_t> nonref;
  p2::PgenReader pgr;
  p2::PreinitPgr(&pgr);
  do {                                                       // the initialisation sequence of PgenReader::Load (pgenlibr.cpp:45-140)
    if (p2::PgfiInitPhase1(path, nullptr, UINT32_MAX, UINT32_MAX, &header_ctrl, &pgfi, &alloc_cacheline_ct, errbuf)) break;
    pgfi.max_allele_ct = 2;
    if ((header_ctrl & 0xc0) == 0xc0) {
      nonref.resize(p2::DivUp(pgfi.raw_variant_ct, p2::kBitsPerWord) + 1);
      pgfi.nonref_flags = nonref.data();
    }
    if (p2::cachealigned_malloc(alloc_cacheline_ct * p2::kCacheline, &pgfi_alloc)) break;
    uint32_t max_vrec_width = 0;
    uintptr_t pgr_alloc_cacheline_ct = 0;
    if (p2::PgfiInitPhase2(header_ctrl, 1, 0, 0, 0, pgfi.raw_variant_ct, &max_vrec_width, &pgfi, pgfi_alloc, &pgr_alloc_cacheline_ct, errbuf)) break;
    if (p2::cachealigned_malloc(pgr_alloc_cacheline_ct * p2::kCacheline, &pgr_alloc)) break;
    if (p2::PgrInit(path, max_vrec_width, &pgfi, &pgr, pgr_alloc)) break;
    if (p2::cachealigned_malloc(p2::DivUp(pgfi.raw_sample_ct, p2::kNypsPerVec) * p2::kBytesPerVec + 64, &genovec)) break;
    rc = (int)p2::PgrValidate(&pgr, genovec, errbuf);
  } while (0);
  if (rc && err && err_len > 0) { strncpy(err, errbuf[0] ? errbuf : "pgenlib initialisation failed", (size_t)err_len - 1); err[err_len - 1] = 0; }
  p2::PglErr e2 = p2::kPglRetSuccess;
  p2::CleanupPgr(&pgr, &e2);
  p2::CleanupPgfi(&pgfi, &e2);
  if (genovec) p2::aligned_free(genovec);
  if (pgfi_alloc) p2::aligned_free(pgfi_alloc);
  if (pgr_alloc) p2::aligned_free(pgr_alloc);
  return rc;
}

}  // extern "C"
