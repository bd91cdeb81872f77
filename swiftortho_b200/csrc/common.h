// Shared declarations of the swiftortho_b200 native library (host side + device launch wrappers).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/swiftortho_b200.h"

namespace so {

typedef long long i64;

// ---- error channel ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

// ---- scoring ---------------------------------------------------------------------------------
// 5-bit residue classes: the 23 letters of the reference's BLOSUM62 dictionary
// (lib/fsearch.py:330) in the order below, both cases; every other byte is class 23 whose row and
// column are -4 (dict2mat default, lib/fsearch.py:333-344).
static const char kB62Letters[] = "ARNDCQEGHILKMFPSTWYVBZX";
enum { kOther = 23, kClasses = 32 };
extern const signed char kB62[23][23];
void make_code_table(uint8_t code[256]);             // byte -> class
void make_score_table(int8_t tbl[kClasses * kClasses]);  // class x class -> score
int score_bytes(uint8_t a, uint8_t b);               // b62[a][b] on raw bytes

// ---- parameters ------------------------------------------------------------------------------
struct Params {
    std::vector<std::string> patterns;                // -s split on ','
    std::vector<std::vector<uint16_t>> alphabets;     // -r split on '/', generate_nr_tbl -> 256-entry table
    int mink = 0, maxk = 0;
    uint32_t nc = 0;
    int step = 1;
    double expect = 1e-3;
    i64 v = 500;
    double max_miss = 1e-3;
    i64 thr = -1;
    bool flt = true;
    i64 chunk = 50000;
    i64 rst = -1, red = -1;
};
int parse_params(const so_params *p, Params &out);

// ---- host algorithms (pure C++; mirrors of reference host-side steps) -------------------------
void seg_mask(const uint8_t *s, i64 n, uint8_t *out);                     // H1
void qsort_perm(const i64 *keys, i64 n, int32_t *perm);                   // Q (full sort)
// Q, pruned: positions [0, need) of the reference quicksort of (key[i], i) — only the partitions
// that intersect [0, need) are refined (the others cannot influence that prefix).
void qsort_prefix(std::vector<uint64_t> &packed, i64 need);
i64 score2bit(i64 raw);
double bit2e(i64 D, i64 ql, i64 tl, i64 bit);
std::string f2s(double e);
std::string fmt_identity(double idy);
int f2s_to(char *out, double e);              // same text into a caller buffer (>= 400 bytes); returns the length
int fmt_identity_to(char *out, double idy);

}  // namespace so
