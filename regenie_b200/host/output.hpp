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

// One prediction file (.loco / .prs, plain or .gz).  The header ids and the first data row are parsed when the file is
// opened (blup_read looks at exactly those, src/Pheno.cpp:1290-1316); the other rows are parsed when a chromosome asks
// for them (blup_read_chr re-reads the file per chromosome, src/Step2_Models.cpp:51-143): a plain file keeps the byte
// offset of every row and holds one parsed row at a time - at N = 500k a .loco file is 120 MB of text, 23 x N tokens -
// a gzip file cannot seek and is parsed once into doubles.
class Loco {
 public:
  std::vector<std::string> ids;                 // sample ids of the header line
  std::vector<double> first;                    // first data row by header position, NaN = NA
  bool prs = false;                             // --use-prs: one row labelled "0", valid for every chromosome
  bool empty() const { return path_.empty(); }  // --ignore-pred
  bool has_row(int chrom) const;
  // values of the row of `chrom` (1..23) by header position, NaN = NA; throws when the file has no such row
  const std::vector<double>& row(int chrom);

 private:
  friend Loco read_loco(const std::string& path, bool prs);
  std::string path_;
  std::vector<int64_t> offs_;                   // [23] byte offset of the row in a plain file, -1 = absent
  std::vector<std::vector<double>> rows_;       // [23] parsed rows (all of them for .gz, the cached one otherwise)
  int cached_ = -1;
  bool eager_ = false;
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

std::string sumstats_header(bool with_info, bool af_cc = false);

// --af-cc (binary traits): allele frequency and sample count among cases / controls, printed after A1FREQ and after N
struct AfCc {
  double af_case = 0, af_control = 0;
  int ns_case = 0, ns_control = 0;
};

// `head` = "CHROM GENPOS ID ALLELE0 ALLELE1 " of the variant; `logp` < 0 or NaN prints NA like a failed test
std::string sumstats_row(const std::string& head, double af, bool with_info, double info, int n, const char* test,
                         double beta, double se, double chisq, double logp, bool test_pass);

// the same row appended to a caller-owned buffer (one buffer per trait, flushed once per block)
void append_sumstats_row(std::string& out, const std::string& head, double af, bool with_info, double info, int n,
                         const char* test, double beta, double se, double chisq, double logp, bool test_pass,
                         const AfCc* cc = nullptr);

// --no-split (print_header_output_all / print_sum_stats_all, src/Step2_Models.cpp:2364-2383, 2441-2493): one file for all
// traits; the variant columns are those of all analysed samples, followed by BETA/SE/CHISQ/LOG10P per trait
std::string sumstats_header_all(int n_pheno, bool with_info = false);
// start of a row: "<head>A1FREQ N N_RR N_RA N_AA TEST"
void append_sumstats_all_start(std::string& out, const std::string& head, double af, int n, long n_rr, long n_ra, long n_aa,
                               const char* test, bool with_info = false, double info = -1.0);
// one trait: " BETA SE CHISQ LOG10P"; `have` false = the trait was ignored for this variant (all NA)
void append_sumstats_all_trait(std::string& out, bool have, double beta, double se, double chisq, double logp, bool test_pass);

// --htp COHORT (HTPv4 rows: print_header_output_htp / print_sum_stats_head_htp / print_sum_stats_htp,
// src/Step2_Models.cpp:2400-2426, :2542-2646) for the single-variant tests of this driver.
struct HtpRow {
  const char* model = "ADD-WGR-LR";   // test + "-WGR" + correction (src/Data.cpp:2075-2102)
  bool bt = false, firth = false;     // trait_mode == 1; --firth
  double beta = 0, se = 0, chisq = 0, logp = 0, af = 0, mac = 0;
  bool test_pass = true;
  long gc[6] = {0, 0, 0, 0, 0, 0};    // cases ref / het / alt, controls ref / het / alt (QT: all samples of the trait in 0-2)
  bool has_score = true;
  double score = 0, skat_var = 0, cal_factor = -1.0;
  double info = -1.0;                 // printed when >= 0 (dosage input)
};
std::string htp_header();
// `head` = "Name\tChr\tPos\tRef\tAlt\t" of the variant
void append_htp_row(std::string& out, const std::string& head, const std::string& trait, const std::string& cohort, const HtpRow& r);

// <out>_<pheno>.regenie.ids: "FID\tIID" of the samples of one trait, no newline after the last one
void write_ids_file(const std::string& path, const std::string& pheno_name, bool print_pheno_name,
                    const std::vector<std::pair<std::string, std::string>>& fid_iid, const uint8_t* mask);

}  // namespace rgh
