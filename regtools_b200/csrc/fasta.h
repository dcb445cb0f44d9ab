// regtools_b200/csrc/fasta.h — reference genome for the intron-motif strand mode.
//
// The reference fetches two 2-mers per junction with faidx (fai_load + fai_fetch per call,
// /root/reference/src/junctions/junctions_extractor.cc:548-584).  Here the FASTA is read once, every sequence is
// kept as the isgraph() bytes of its lines — exactly what fai_fetch returns (htslib faidx.c:341-415, index rules
// faidx.c:82-155) — and the whole genome lives in HBM as one byte per base (case and IUPAC codes preserved: a
// lower-case or N motif must stay unmatched, as in the reference).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rtjx {

struct FastaGenome {
    std::vector<std::string> names;       // header up to the first white space; a repeated name is ignored (first wins)
    std::vector<uint64_t> offset;         // start of the sequence in `bases`
    std::vector<uint64_t> length;
    std::vector<uint8_t> bases;           // all sequences back to back (+16 bytes of padding)
    int find(const std::string& name) const;
};

// false + *err on failure (unreadable, compressed, no sequence).
bool load_fasta(const std::string& path, FastaGenome* out, std::string* err);

}  // namespace rtjx
