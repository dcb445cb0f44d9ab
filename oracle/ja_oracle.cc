// oracle/ja_oracle.cc — TEST INFRASTRUCTURE ONLY.  NOT part of the product (see jx_oracle.h for the rules).
//
// CPU restatement of `regtools junctions annotate`, single-threaded and literal:
//   driver            /root/reference/src/junctions/junctions_main.cc:61-92
//   annotator         /root/reference/src/junctions/junctions_annotator.cc  (:66-81 ends, :94-114 splice site,
//                     :128-213 overlap_ps, :246-311 overlap_ns, :314-328 anchor, :333-363 check_for_overlap,
//                     :367-388 bin walk), junctions_annotator.h:69-121 (header + print)
//   GTF               /root/reference/src/gtf/gtf_parser.cc (:62-87 exon line, :90-106 attributes, :109-125 transcript map,
//                     :149-169 bins, :192-208 exon order)
//   BED reader        /root/reference/src/utils/bedtools/bedFile (GetHeader, GetNextBed, parseLine; bins bedFile.h:49-63)
//   FASTA             htslib faidx.c:341-415 (fai_fetch clipping), common.h:59-83 (rev_comp)
// C++ (not C) because the outputs are std::set<std::string> orders and std::sort's treatment of equal exon starts: both
// are properties of libstdc++ that the restatement uses directly instead of imitating.
//
// Parity status: PINNED — tests/test_annotate_cpu.py checks it against the reference's own golden
// (tests/integration-test/data/junctions-annotate/expected-annotate.out, copied to tests/golden/annotate/) and against
// outputs of the UNMODIFIED reference (oracle/_ref/regtools_ref_annotate) on generated GTF/BED/FASTA fixtures.
//
// Stated divergences (the reference has undefined behaviour there): exons[i + 1] past the last exon (junctions_annotator.cc
// :145,:263) reads beyond the vector — here it never matches; an attribute without a value (gtf_parser.cc:100-102) is skipped.
//
//   ja_oracle [-S] [-o out.tsv] junctions.bed ref.fa annotations.gtf         exit 0 / 1 like the reference
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unistd.h>
#include <vector>

using std::string;
using std::vector;

static vector<string> split(const string& s, char d) {          // lineFileUtilities.h Tokenize: getline semantics
    vector<string> out;
    std::stringstream ss(s);
    string item;
    while (std::getline(ss, item, d)) out.push_back(item);
    return out;
}

struct Exon { uint32_t start, end; string chrom, strand; };
struct Transcript { vector<Exon> exons; string gene_name, gene_id; bool has_gene = false; };

static const uint32_t BIN_OFFSETS[7] = {32678 + 4096 + 512 + 64 + 8 + 1, 4096 + 512 + 64 + 8 + 1, 512 + 64 + 8 + 1, 64 + 8 + 1, 8 + 1, 1, 0};

static uint32_t get_bin(uint32_t start, uint32_t end) {          // bedFile.h getBin
    --end;
    start >>= 14; end >>= 14;
    for (int i = 0; i < 7; ++i) {
        if (start == end) return BIN_OFFSETS[i] + start;
        start >>= 3; end >>= 3;
    }
    return 0;
}

struct Gtf {
    std::map<string, Transcript> tx;
    std::map<string, std::map<uint32_t, vector<string> > > chrbin;

    static string attribute(const vector<string>& attrs, const string& name) {     // gtf_parser.cc:90-106
        for (size_t i = 0; i < attrs.size(); ++i) {
            string a = attrs[i];
            if (!a.empty() && a[0] == ' ') a.erase(0, 1);
            vector<string> tok = split(a, ' ');
            if (tok.size() < 2) continue;                                           // (UB in the reference)
            if (tok[0] == name) {
                string v = tok[1];
                if (!v.empty() && v[0] == '"' && v[v.size() - 1] == '"') { v.erase(v.begin()); v.erase(v.end() - 1); }
                return v;
            }
        }
        return "NA";
    }
    void load(const string& path) {
        std::ifstream f(path.c_str());
        if (!f.is_open()) { std::cerr << "\nUnable to open GTF file."; exit(1); }
        string line;
        while (std::getline(f, line)) {
            if (line.empty()) throw std::runtime_error("Expected 9 fields in GTF line.");   // the reference dies in line.at(0)
            if (line[0] == '#') continue;
            vector<string> fld = split(line, '\t');
            if (fld.size() != 9) { std::cerr << line << std::endl << fld.size(); throw std::runtime_error("Expected 9 fields in GTF line."); }
            if (fld[2] != "exon") continue;
            Exon e;
            e.chrom = fld[0]; e.start = (uint32_t)atol(fld[3].c_str()); e.end = (uint32_t)atol(fld[4].c_str()); e.strand = fld[6];
            vector<string> attrs = split(fld[8], ';');
            string tid = attribute(attrs, "transcript_id"), gname = attribute(attrs, "gene_name"), gid = attribute(attrs, "gene_id");
            if (tid == "NA") continue;
            Transcript& t = tx[tid];
            t.exons.push_back(e);
            if (!t.has_gene) { t.gene_name = gname; t.gene_id = gid; t.has_gene = true; }
        }
        // GtfParser::load (:262-268) sorts TWICE: construct_junctions() sorts because transcripts_sorted_ is still false, then
        // load() calls sort_exons_within_transcripts() again.  Each pass picks its direction from exons[0].strand AT THAT
        // MOMENT (:192-208), so a transcript whose exons disagree on the strand can be re-sorted the other way round.
        for (int pass = 0; pass < 2; ++pass)
            for (std::map<string, Transcript>::iterator it = tx.begin(); it != tx.end(); ++it) {
                vector<Exon>& ex = it->second.exons;
                if (ex[0].strand == "+") std::sort(ex.begin(), ex.end(), [](const Exon& a, const Exon& b) { return a.start < b.start; });
                else if (ex[0].strand == "-") std::sort(ex.begin(), ex.end(), [](const Exon& a, const Exon& b) { return a.start > b.start; });
                else { std::cerr << "Undefined strand for exon " << ex[0].start << ex[0].end; exit(1); }
            }
        for (std::map<string, Transcript>::iterator it = tx.begin(); it != tx.end(); ++it) {      // :149-169
            const vector<Exon>& ex = it->second.exons;
            chrbin[ex[0].chrom][get_bin(ex[0].start, ex[ex.size() - 1].end)].push_back(it->first);
        }
    }
};

struct Fasta {                                                     // as faidx sees the file (faidx.c:82-155)
    std::map<string, string> seq;
    bool load(const string& path) {
        std::ifstream f(path.c_str());
        if (!f.is_open()) return false;
        string line, name; bool keep = false;
        while (std::getline(f, line)) {
            if (!line.empty() && line[0] == '>') {
                size_t e = 1;
                while (e < line.size() && !isspace((unsigned char)line[e])) ++e;
                name = line.substr(1, e - 1);
                keep = !seq.count(name);
                if (keep) seq[name] = "";
            } else if (keep) {
                string& s = seq[name];
                for (size_t i = 0; i < line.size(); ++i) if (isgraph((unsigned char)line[i])) s.push_back(line[i]);
            }
        }
        return true;
    }
    // fai_fetch("chrom:b1-e1") (faidx.c:341-415): b1 / e1 are the numbers the annotator printed with uint32 arithmetic; faidx
    // reads them back with atoi into ints (4294967295 -> -1), decrements a positive beg, clips both to the length and never
    // clamps a negative beg: it then starts reading that many bytes BEFORE the sequence — for -1 the header's newline, which is
    // skipped as non-graph, so the result is the first (end - beg) bases.  (beg < -1 would read header text; treated as -1.)
    bool fetch(const string& chrom, uint32_t b1, uint32_t e1, string* out) const {
        std::map<string, string>::const_iterator it = seq.find(chrom);
        if (it == seq.end()) return false;
        long long len = (long long)it->second.size(), beg = (int32_t)b1, end = (int32_t)e1;
        if (beg > 0) --beg;
        if (beg >= len) beg = len;
        if (end >= len) end = len;
        if (beg > end) beg = end;
        const long long n = end - beg, from = beg < 0 ? 0 : beg;
        *out = it->second.substr((size_t)from, (size_t)std::min<long long>(n, len - from));
        return true;
    }
};

static string rev_comp(const string& s) {                          // common.h:59-83
    string rc;
    for (int i = (int)s.size() - 1; i >= 0; --i) {
        char c = s[i];
        rc.push_back(c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N');
    }
    return rc;
}

struct AJ {
    string chrom, name, score, strand, splice_site, anchor;
    uint32_t start, end;
    bool kd, ka, kj;
    std::set<string> transcripts, exons_skipped;
    std::set<vector<string> > genes;
    std::set<uint32_t> acceptors_skipped, donors_skipped;
    void reset() { anchor = "N"; splice_site = ""; kd = ka = kj = false; transcripts.clear(); exons_skipped.clear(); genes.clear(); acceptors_skipped.clear(); donors_skipped.clear(); }
};

static string u2s(uint32_t v) { std::stringstream s; s << v; return s.str(); }

static void annotate_anchor(AJ& j) {                               // :314-328
    j.anchor = "N";
    if (j.kj) j.anchor = "DA";
    else if (j.kd) j.anchor = j.ka ? "NDA" : "D";
    else if (j.ka) j.anchor = "A";
}

static bool overlap_ps(const vector<Exon>& ex, AJ& j, bool skip_single) {          // :128-213
    if (skip_single && ex.size() == 1) return false;
    bool started = false;
    if (ex[0].start > j.end || ex[ex.size() - 1].end < j.start) return false;
    for (size_t i = 0; i < ex.size(); ++i) {
        if (ex[i].start > j.end) break;
        if (ex[i].end == j.start && i + 1 < ex.size() && ex[i + 1].start == j.end) {
            j.ka = j.kd = j.kj = true;
        } else {
            if (!started && ex[i].end >= j.start) started = true;
            if (started) {
                if (ex[i].start > j.start && ex[i].end < j.end && i > 0 && i < ex.size() - 1) j.exons_skipped.insert(u2s(ex[i].start) + "-" + u2s(ex[i].end));
                if (ex[i].end > j.start && ex[i].end < j.end && i < ex.size() - 1) j.donors_skipped.insert(ex[i].end);
                if (ex[i].start < j.end && ex[i].start > j.start && i > 0) j.acceptors_skipped.insert(ex[i].start);
                if (ex[i].end == j.start) j.kd = true;
                if (ex[i].start == j.end) j.ka = true;
            }
        }
    }
    annotate_anchor(j);
    return j.anchor != "N";
}

static bool overlap_ns(const vector<Exon>& ex, AJ& j, bool skip_single) {          // :246-311
    if (skip_single && ex.size() == 1) return false;
    bool started = false;
    if (ex[0].end < j.start || ex[ex.size() - 1].start > j.end) return false;
    for (size_t i = 0; i < ex.size(); ++i) {
        if (ex[i].end < j.start) break;
        if (ex[i].start == j.end && i + 1 < ex.size() && ex[i + 1].end == j.start) {
            j.ka = j.kd = j.kj = true;
        } else {
            if (!started && ex[i].start <= j.end) started = true;
            if (started) {
                if (ex[i].start > j.start && ex[i].end < j.end && i > 0 && i < ex.size() - 1) j.exons_skipped.insert(u2s(ex[i].start) + "-" + u2s(ex[i].end));
                if (ex[i].end > j.start && ex[i].end < j.end && i < ex.size() - 1) j.acceptors_skipped.insert(ex[i].end);
                if (ex[i].start < j.end && ex[i].start > j.start) j.donors_skipped.insert(ex[i].start);
                if (ex[i].end == j.start) j.ka = true;
                if (ex[i].start == j.end) j.kd = true;
            }
        }
    }
    annotate_anchor(j);
    return j.anchor != "N";
}

static void annotate(Gtf& g, AJ& j, bool skip_single) {            // :333-388
    uint32_t sb = j.start >> 14, eb = (j.end - 1) >> 14;
    for (int lvl = 0; lvl < 7; ++lvl) {
        const uint32_t off = BIN_OFFSETS[lvl];
        for (uint32_t b = sb + off; b <= eb + off; ++b) {
            const vector<string> ids = g.chrbin[j.chrom][b];
            for (size_t k = 0; k < ids.size(); ++k) {
                const Transcript& t = g.tx[ids[k]];
                if (j.strand != t.exons[0].strand) continue;
                bool hit;
                if (j.strand == "+") hit = overlap_ps(t.exons, j, skip_single);
                else if (j.strand == "-") hit = overlap_ns(t.exons, j, skip_single);
                else throw std::runtime_error("Unknown strand " + j.strand + "\n\n");
                if (hit) {
                    j.transcripts.insert(ids[k]);
                    vector<string> gv; gv.push_back(t.gene_name); gv.push_back(t.gene_id);
                    j.genes.insert(gv);
                }
            }
            if (b == 0xffffffffu) break;
        }
        sb >>= 3; eb >>= 3;
    }
}

static bool is_integer(const string& s) {
    if (s.empty()) return false;
    char* e = NULL;
    strtol(s.c_str(), &e, 10);
    return *e == '\0';
}

int main(int argc, char** argv) {
    bool skip_single = true; string out_path = "NA";
    int c;
    while ((c = getopt(argc, argv, "So:")) != -1) {
        if (c == 'S') skip_single = false;
        else if (c == 'o') out_path = optarg;
        else { std::cerr << "Error parsing inputs!(1)\n\n" << std::endl; return 1; }
    }
    if (argc - optind != 3) { std::cerr << "Error parsing inputs!(2)\n\n" << std::endl; return 1; }
    const string bed = argv[optind], fa = argv[optind + 1], gtf = argv[optind + 2];
    std::ofstream ofs;
    try {
        Gtf g;
        g.load(gtf);
        std::ifstream bf(bed.c_str());
        if (!bf.is_open()) { std::cerr << "Error: The requested file (" << bed << ") could not be opened." << std::endl; return 1; }
        if (out_path != "NA") { ofs.open(out_path.c_str()); if (!ofs.is_open()) throw std::runtime_error("Unable to open " + out_path); }
        std::ostream& out = out_path == "NA" ? std::cout : ofs;
        out << "chrom\tstart\tend\tname\tscore\tstrand\tsplice_site\tacceptors_skipped\texons_skipped\tdonors_skipped\tanchor"
               "\tknown_donor\tknown_acceptor\tknown_junction\tgene_names\tgene_ids\ttranscripts\n";
        Fasta fasta; bool fasta_loaded = false, fasta_ok = false;
        string line; bool header = true; size_t n_fields0 = 0; int linec = 0;
        AJ j;
        // (the BED reader's errors are exit(1) inside bedFile.h: automatic objects are not destroyed, so an ofstream that
        // only holds the header so far is never flushed — the file stays empty)
        while (std::getline(bf, line)) {
            if (header && (line.find("#") == 0 || line.find("browser") == 0 || line.find("track") == 0)) continue;   // GetHeader
            if (header && bf.eof()) break;         // GetHeader's own getline hit EOF: the stream is not good() for GetNextBed (bedFile.cpp:208-214)
            header = false;
            if (!line.empty() && line[line.size() - 1] == '\r') line.resize(line.size() - 1);
            vector<string> f = split(line, '\t');
            if (f.empty()) break;                                                  // BED_BLANK ends the loop
            if (f[0].find("#") == 0 || f[0].find("browser") == 0 || f[0].find("track") == 0) break;   // BED_HEADER too
            if (f.size() < 3) { std::cerr << "It looks as though you have less than 3 columns" << std::endl; exit(1); }
            if (!is_integer(f[1]) || !is_integer(f[2])) { std::cerr << "Unexpected file format." << std::endl; exit(1); }
            if (!n_fields0) n_fields0 = f.size();
            if (f.size() != n_fields0) { std::cerr << "Differing number of BED fields encountered" << std::endl; exit(1); }
            j.reset();
            {   // parseBedLine, bedFile.h:685-760
                const int is = atoi(f[1].c_str()), ie = atoi(f[2].c_str());
                if (is < 0) { std::cerr << "Error: malformed BED entry. Start Coordinate detected that is < 0. Exiting." << std::endl; exit(1); }
                if (ie < 0) { std::cerr << "Error: malformed BED entry. End Coordinate detected that is < 0. Exiting." << std::endl; exit(1); }
                j.start = (uint32_t)is; j.end = (uint32_t)ie;
                if (j.start == j.end) { j.start--; j.end++; }
                if (j.start > j.end) { std::cerr << "Error: malformed BED entry. Start was greater than end. Exiting." << std::endl; exit(1); }
            }
            j.chrom = f[0];
            j.name = f.size() > 3 ? f[3] : ""; j.score = f.size() > 4 ? f[4] : ""; j.strand = f.size() > 5 ? f[5] : "";
            if (f.size() != 12 || f[10].empty()) throw std::runtime_error("BED line not in BED12 format. start: " + j.chrom + ":" + u2s(j.start));
            vector<string> bs = split(f[10], ',');                                 // adjust_junction_ends :66-81
            j.start += (uint32_t)atoi(bs[0].c_str());
            j.end -= (uint32_t)(atoi(bs.size() > 1 ? bs[1].c_str() : "0") - 1);
            if (!fasta_loaded) { fasta_ok = fasta.load(fa); fasta_loaded = true; }
            string s1, s2;                                                         // get_splice_site :94-114
            const string p1 = j.chrom + ":" + u2s(j.start + 1) + "-" + u2s(j.start + 2), p2 = j.chrom + ":" + u2s(j.end - 2) + "-" + u2s(j.end - 1);
            std::cerr << "position = " << p1 << std::endl;
            if (!fasta_ok || !fasta.fetch(j.chrom, j.start + 1u, j.start + 2u, &s1))
                throw std::runtime_error("Unable to extract FASTA sequence for position " + p1 + "\n\n");
            std::cerr << "position = " << p2 << std::endl;
            if (!fasta.fetch(j.chrom, j.end - 2u, j.end - 1u, &s2))
                throw std::runtime_error("Unable to extract FASTA sequence for position " + p2 + "\n\n");
            if (j.strand == "-") j.splice_site = rev_comp(s2) + "-" + rev_comp(s1);
            else j.splice_site = s1 + "-" + s2;
            annotate(g, j, skip_single);
            out << j.chrom << "\t" << j.start << "\t" << j.end << "\t" << j.name << "\t" << j.score << "\t" << j.strand << "\t" << j.splice_site
                << "\t" << j.acceptors_skipped.size() << "\t" << j.exons_skipped.size() << "\t" << j.donors_skipped.size() << "\t" << j.anchor
                << "\t" << j.kd << "\t" << j.ka << "\t" << j.kj;
            if (!j.genes.empty()) {
                out << "\t";
                for (std::set<vector<string> >::iterator it = j.genes.begin(); it != j.genes.end(); ++it) out << (it != j.genes.begin() ? "," : "") << (*it)[0];
                out << "\t";
                for (std::set<vector<string> >::iterator it = j.genes.begin(); it != j.genes.end(); ++it) out << (it != j.genes.begin() ? "," : "") << (*it)[1];
            } else out << "\tNA\tNA";
            if (!j.transcripts.empty()) {
                out << "\t";
                for (std::set<string>::iterator it = j.transcripts.begin(); it != j.transcripts.end(); ++it) out << (it != j.transcripts.begin() ? "," : "") << *it;
            } else out << "\tNA";
            out << std::endl;
            ++linec;
        }
        std::cerr << std::endl << "Annotated " << linec << " lines." << std::endl;
    } catch (const std::runtime_error& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
