// PLINK 2 .pgen / .pvar / .psam reader of the rgb200 host driver: biallelic hard calls, storage modes 0x02
// (fixed-width 2-bit) and 0x10 (variable-width records: plain, 1-bit + difflist, difflist over a constant,
// LD-compressed against the previous plain record).  The reference reads these through the vendored pgenlib
// (ReadHardcalls(..., allele_idx = 1), src/Geno.cpp:1773-1821; .pvar / .psam parsing :771-1011); the layout is
// restated from the format pgenlib documents (external_libs/pgenlib/include/pgenlib_read.cc).  Every variant is
// handed to the GPU as a PLINK 1 2-bit row (ALT count 0/1/2/missing -> codes 11/10/00/01), so the .bed kernels
// and entry points serve .pgen input unchanged.
#pragma once
#include "data.hpp"

namespace rgh {

// The records a block of variants needs, sliced out of the file for rg_pgen_decode (the device expands them; the
// reference's counterpart is one PgenReader::Read per variant, src/Geno.cpp:1773-1821 / :2538-2594).
struct PgenBatch {
  std::vector<uint8_t> bytes;                   // records, each starting at a multiple of 16 bytes
  std::vector<uint64_t> rec_off;
  std::vector<uint32_t> rec_len;
  std::vector<uint8_t> rec_type;                // low 3 bits of the variant record type
  std::vector<int32_t> own, base;               // per variant: its record; the non-LD record an LD-compressed one refers to
};

struct PgenFile {
  std::string prefix;
  int sex_specific = 0;                         // 1 = males only, 2 = females only (set before open)
  std::vector<Snp> snps;                        // after filters; offset = variant index in the .pgen
  std::vector<std::string> keys_file, keys;
  std::vector<std::pair<std::string, std::string>> ids_file;   // (FID, IID) in .psam order
  std::vector<int> sex_file;
  std::vector<int32_t> sample_idx;
  std::map<std::string, uint32_t> key_to_ind;
  uint32_t n_file = 0, m_file = 0;
  uint64_t row_stride = 0;                      // bytes per emitted .bed-coded row
  std::vector<uint8_t> data;                    // whole .pgen
  std::vector<uint8_t> vrtype;
  std::vector<uint64_t> fpos;                   // m_file + 1 record offsets
  void open(const std::string& prefix, const std::set<std::string>& exclude, const std::set<std::string>& extract,
            const std::set<std::string>& remove, const std::set<std::string>& keep, const std::set<int>& chrs);
  // rows of snps[first .. first+n) as PLINK 1 2-bit rows (ref-last coding: code 00 = two copies of ALT)
  void read_rows(size_t first, size_t n, uint8_t* out);
  // the same variants as record bytes + indices (no decode on the host)
  void gather(size_t first, size_t n, PgenBatch& out) const;

 private:
  long base_idx_ = -1;
  std::vector<uint8_t> base_, cur_;             // one value (0..3) per sample
  void decode_nonld(uint32_t v, std::vector<uint8_t>& g) const;
  void decode(uint32_t v);
  size_t difflist(size_t p, std::vector<uint32_t>& ids, std::vector<uint8_t>& vals) const;
};

}  // namespace rgh
