// H0: FASTA container with the reference's record semantics (lib/fsearch.py:1543-1553 `index`,
// 2180-2202 `Fasta.__getitem__`): record starts are offset 0 plus every '>' that follows '\n';
// header = first line minus its first byte; sequence = the other lines joined, bytes untouched.
// The product layout is one packed residue buffer + offsets[n+1], ready for a single H2D copy.
#include <string>
#include <vector>

#include "common.h"

struct so_fasta {
    std::string text;
    std::vector<uint8_t> residues;
    std::vector<uint64_t> offsets;          // n+1
    std::vector<uint64_t> hd_start, hd_len;  // into text
};

extern "C" {

int so_fasta_open(const char *path, so_fasta **out) {
    if (!path || !out) {
        so::set_error("so_fasta_open: null argument");
        return SO_EINVAL;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        so::set_error("cannot open %s", path);
        return SO_EIO;
    }
    so_fasta *F = new so_fasta();
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    F->text.resize((size_t)sz);
    if (sz > 0 && fread(&F->text[0], 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        delete F;
        so::set_error("short read on %s", path);
        return SO_EIO;
    }
    fclose(f);
    const std::string &t = F->text;
    std::vector<size_t> starts;
    starts.push_back(0);
    for (size_t i = 1; i < t.size(); i++)
        if (t[i] == '>' && t[i - 1] == '\n') starts.push_back(i);
    F->residues.reserve(t.size());
    F->offsets.push_back(0);
    for (size_t r = 0; r < starts.size(); r++) {
        size_t b = starts[r], e = r + 1 < starts.size() ? starts[r + 1] : t.size();
        size_t nl = t.find('\n', b);
        if (nl == std::string::npos || nl >= e) nl = e;
        F->hd_start.push_back(b < e ? b + 1 : b);
        F->hd_len.push_back(nl > b ? nl - b - 1 : 0);
        for (size_t i = nl + 1; i < e; i++)
            if (t[i] != '\n') F->residues.push_back((uint8_t)t[i]);
        F->offsets.push_back(F->residues.size());
    }
    *out = F;
    return SO_OK;
}

void so_fasta_close(so_fasta *f) { delete f; }

int64_t so_fasta_count(const so_fasta *f) { return f ? (int64_t)f->offsets.size() - 1 : 0; }

int64_t so_fasta_residues(const so_fasta *f, const uint8_t **residues, const uint64_t **offsets) {
    if (!f) return 0;
    if (residues) *residues = f->residues.data();
    if (offsets) *offsets = f->offsets.data();
    return (int64_t)f->residues.size();
}

int so_fasta_header(const so_fasta *f, int64_t i, const char **hd, int64_t *len) {
    if (!f || i < 0 || i >= so_fasta_count(f)) {
        so::set_error("so_fasta_header: record out of range");
        return SO_EINVAL;
    }
    *hd = f->text.data() + f->hd_start[(size_t)i];
    *len = (int64_t)f->hd_len[(size_t)i];
    return SO_OK;
}
}
