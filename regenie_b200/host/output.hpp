// Text formats either side of the hot path: the `.loco` / `.prs` prediction files Step 1 writes and Step 2 reads,
// `_pred.list`, the `.regenie` summary-statistics rows and the `.regenie.ids` sample lists.
// Restated from the reference (not copied):
//   write_predictions / write_ID_header / write_chr_row   src/Data.cpp:1795-1982
//   blup_read / blup_read_chr                             src/Pheno.cpp:1238-1345, src/Step2_Models.cpp:51-143
//   check_blup (pred.list)                                src/Pheno.cpp:1204-1229
//   print_header_output_single / print_sum_stats_single   src/Step2_Models.cpp:2386-2398, 2502-2540
//   get_logp                                              src/Regenie.cpp:1843-1857
//   write_ids                                             src/Pheno.cpp:1538-1576
#pragma once
#include "textio.hpp"

namespace rgh {

// one prediction file: sample ids of the header line and the token rows by chromosome
struct Loco {
  std::vector<std::string> ids;
  std::vector<std::vector<std::string>> rows;   // [23] for .loco; rows[0] = the single "0" row for .prs
  bool prs = false;
};

// `prs`: the file is a whole-genome PRS (--use-prs): one row whose first token must be "0"
Loco read_loco(const std::string& path, bool prs = false);

// pred.list / prs.list: "<phenotype> <file>" per line
std::map<std::string, std::string> read_pred_list(const std::string& path);

// header line "FID_IID id1 id2 ... \n" followed by one line per entry of `row_labels`; values[r] points to N doubles
// indexed by sample, `order` lists the samples to print (std::map key order of FID_IID, analysed samples only),
// masked samples print NA; every token is followed by a space and numbers print like `ostream << double`
void write_pred_file(TextWriter& out, const std::vector<std::string>& keys, const std::vector<uint32_t>& order,
                     const uint8_t* mask, const std::vector<int>& row_labels, const std::vector<const double*>& values);

// -log10 of the chi-square(1) tail probability of `t`, with the reference's underflow fallback
double get_logp(double t);

std::string sumstats_header(bool with_info);

// `head` = "CHROM GENPOS ID ALLELE0 ALLELE1 " of the variant; `logp` < 0 or NaN prints NA like a failed test
std::string sumstats_row(const std::string& head, double af, bool with_info, double info, int n, const char* test,
                         double beta, double se, double chisq, double logp, bool test_pass);

// <out>_<pheno>.regenie.ids: "FID\tIID" of the samples of one trait, no newline after the last one
void write_ids_file(const std::string& path, const std::string& pheno_name, bool print_pheno_name,
                    const std::vector<std::pair<std::string, std::string>>& fid_iid, const uint8_t* mask);

}  // namespace rgh
