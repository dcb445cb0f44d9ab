// regtools_b200/csrc/annotate.cc — host side of `junctions annotate` on the B200: GTF and BED12 readers
// that keep the reference's parsing rules, flat arrays for the device, the launch, and the TSV writer.
//
// Reference: driver src/junctions/junctions_main.cc:61-92; JunctionsAnnotator src/junctions/junctions_annotator.{h,cc};
// GtfParser src/gtf/gtf_parser.cc; BedFile src/utils/bedtools/bedFile (GetHeader / GetNextBed / parseLine).
// There is no CPU path for the annotation itself: without a CUDA device rtjx_annotate fails with RTJX_E_CUDA.
#include "../../include/rtjx.h"
#include "annotate.cuh"
#include "fasta.h"

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace rtjx {
namespace {

std::vector<std::string> split(const std::string& s, char d) {              // Tokenize (lineFileUtilities.h): getline rules
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, d)) out.push_back(item);
    return out;
}

const uint32_t BIN_OFFSETS[7] = {32678 + 4096 + 512 + 64 + 8 + 1, 4096 + 512 + 64 + 8 + 1, 512 + 64 + 8 + 1, 64 + 8 + 1, 8 + 1, 1, 0};
uint32_t get_bin(uint32_t start, uint32_t end) {                              // bedFile.h getBin
    --end;
    start >>= 14; end >>= 14;
    for (int i = 0; i < 7; ++i) {
        if (start == end) return BIN_OFFSETS[i] + start;
        start >>= 3; end >>= 3;
    }
    return 0;
}

struct Status { int code = RTJX_OK; std::string msg; bool ok() const { return code == RTJX_OK; } };
Status fail(int code, const std::string& m) { Status s; s.code = code; s.msg = m; return s; }

// ---- GTF (gtf_parser.cc) ------------------------------------------------------------------------------------
struct ExonRec { uint32_t start, end; uint32_t chrom; uint8_t strand; };     // strand: 0 '+', 1 '-', 2 other
struct GtfFlat {
    std::vector<std::string> tx_id, gene_name, gene_id;                       // per transcript, id order
    std::vector<uint32_t> tx_ex_off, ex_start, ex_end, bin_off, bin_tx;
    std::vector<uint8_t> tx_strand;
    std::vector<unsigned long long> bin_key;
    std::map<std::string, uint32_t> chrom_id;
};

std::string attribute(const std::vector<std::string>& attrs, const char* name) {   // parse_attribute :90-106
    for (size_t i = 0; i < attrs.size(); ++i) {
        std::string a = attrs[i];
        if (!a.empty() && a[0] == ' ') a.erase(0, 1);
        const std::vector<std::string> tok = split(a, ' ');
        if (tok.size() < 2) continue;                                         // (the reference indexes tokens[1] regardless)
        if (tok[0] == name) {
            std::string v = tok[1];
            if (!v.empty() && v[0] == '"' && v[v.size() - 1] == '"') { v.erase(v.begin()); v.erase(v.end() - 1); }   // common::unquote
            return v;
        }
    }
    return "NA";
}

Status load_gtf(const std::string& path, GtfFlat* out) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) return fail(RTJX_E_IO, "\nUnable to open GTF file.");
    struct Tx { std::vector<ExonRec> exons; std::string gene_name, gene_id; bool has_gene = false; };
    std::map<std::string, Tx> txs;                                            // transcript_map_: std::string order
    std::vector<std::string> chroms;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) return fail(RTJX_E_IO, "Expected 9 fields in GTF line.");   // the reference dies in line.at(0)
        if (line[0] == '#') continue;
        const std::vector<std::string> fld = split(line, '\t');
        if (fld.size() != 9) return fail(RTJX_E_IO, "Expected 9 fields in GTF line.");
        if (fld[2] != "exon") continue;
        const std::vector<std::string> attrs = split(fld[8], ';');
        const std::string tid = attribute(attrs, "transcript_id");
        if (tid == "NA") continue;
        ExonRec e;
        e.start = (uint32_t)atol(fld[3].c_str()); e.end = (uint32_t)atol(fld[4].c_str());
        e.strand = fld[6] == "+" ? 0 : (fld[6] == "-" ? 1 : 2);
        std::map<std::string, uint32_t>::iterator ci = out->chrom_id.find(fld[0]);
        if (ci == out->chrom_id.end()) ci = out->chrom_id.insert(std::make_pair(fld[0], (uint32_t)out->chrom_id.size())).first;
        e.chrom = ci->second;
        Tx& t = txs[tid];
        t.exons.push_back(e);
        if (!t.has_gene) { t.gene_name = attribute(attrs, "gene_name"); t.gene_id = attribute(attrs, "gene_id"); t.has_gene = true; }
    }
    // sort_exons_within_transcripts (:192-208): by the strand of the FIRST exon in file order; same std::sort, same input
    // order, same comparator results as the reference => the same arrangement of equal starts
    // GtfParser::load (:262-268) sorts twice (construct_junctions() sorts first because transcripts_sorted_ is still false);
    // each pass takes its direction from exons[0].strand at that moment, which matters for transcripts with mixed strands
    for (int pass = 0; pass < 2; ++pass)
        for (std::map<std::string, Tx>::iterator it = txs.begin(); it != txs.end(); ++it) {
            std::vector<ExonRec>& ex = it->second.exons;
            if (ex[0].strand == 0) std::sort(ex.begin(), ex.end(), [](const ExonRec& a, const ExonRec& b) { return a.start < b.start; });
            else if (ex[0].strand == 1) std::sort(ex.begin(), ex.end(), [](const ExonRec& a, const ExonRec& b) { return a.start > b.start; });
            else return fail(RTJX_E_IO, "Undefined strand for exon " + std::to_string(ex[0].start) + std::to_string(ex[0].end));
        }
    // annotate_transcript_with_bins (:149-169) + flat arrays; transcripts are numbered in map (id) order
    std::map<unsigned long long, std::vector<uint32_t> > bins;
    out->tx_ex_off.push_back(0);
    uint32_t t_index = 0;
    for (std::map<std::string, Tx>::iterator it = txs.begin(); it != txs.end(); ++it, ++t_index) {
        const std::vector<ExonRec>& ex = it->second.exons;
        out->tx_id.push_back(it->first); out->gene_name.push_back(it->second.gene_name); out->gene_id.push_back(it->second.gene_id);
        out->tx_strand.push_back(ex[0].strand);
        for (size_t k = 0; k < ex.size(); ++k) { out->ex_start.push_back(ex[k].start); out->ex_end.push_back(ex[k].end); }
        out->tx_ex_off.push_back((uint32_t)out->ex_start.size());
        bins[(unsigned long long)ex[0].chrom << 32 | get_bin(ex[0].start, ex[ex.size() - 1].end)].push_back(t_index);
    }
    out->bin_off.push_back(0);
    for (std::map<unsigned long long, std::vector<uint32_t> >::iterator it = bins.begin(); it != bins.end(); ++it) {
        out->bin_key.push_back(it->first);
        out->bin_tx.insert(out->bin_tx.end(), it->second.begin(), it->second.end());
        out->bin_off.push_back((uint32_t)out->bin_tx.size());
    }
    return Status();
}

// ---- BED12 junctions (BedFile::GetHeader / GetNextBed / parseLine, adjust_junction_ends :66-81) ---------------
// The file is read into one buffer and tokenised in place (tabs become NULs); the text columns the writer needs stay as
// views into that buffer.  Same rules as the reference's getline-based Tokenize: consecutive tabs give empty fields, a
// trailing tab gives none, an empty line gives no field at all.
struct Tok { const char* p; uint32_t n; };
struct JunctionLines {
    std::string buf;                                                          // the whole file
    std::vector<Tok> chrom, name, score, strand;
    std::vector<uint32_t> start, end;                                         // adjusted
    Status stop;                                                              // why reading stopped early (error after the lines above)
    // every `stop` of read_junctions except "not BED12" is an exit(1) inside the reference's BedFile: its ofstream is never
    // destroyed, so a header that no junction line (std::endl) has flushed yet never reaches the -o file
    bool stop_is_exit = false;
};

bool is_integer(const char* s) {                                              // s is NUL-terminated
    if (!*s) return false;
    char* e = nullptr;
    strtol(s, &e, 10);
    return *e == '\0';
}
bool is_header(const char* s, size_t n) {
    return (n >= 1 && s[0] == '#') || (n >= 7 && memcmp(s, "browser", 7) == 0) || (n >= 5 && memcmp(s, "track", 5) == 0);
}

Status read_junctions(const std::string& path, JunctionLines* out) {
    {
        const bool from_stdin = path == "stdin" || path == "-";             // BedFile::Open (bedFile.cpp:99-101)
        FILE* f = from_stdin ? stdin : fopen(path.c_str(), "rb");
        if (!f) return fail(RTJX_E_IO, "Error: The requested file (" + path + ") could not be opened. Exiting!");
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) out->buf.append(chunk, got);
        if (!from_stdin) fclose(f);
        if (out->buf.size() >= 2 && (unsigned char)out->buf[0] == 0x1f && (unsigned char)out->buf[1] == 0x8b)
            return fail(RTJX_E_UNSUPPORTED, "gzip-compressed junction files are not read by the B200 path: " + path);
    }
    std::string& b = out->buf;
    b.push_back('\n');                                                        // sentinel: every line ends in the buffer
    const size_t total = b.size() - 1;
    {
        size_t lines = 0;
        for (size_t k = 0; k < total; ++k) lines += b[k] == '\n';
        out->chrom.reserve(lines + 1); out->name.reserve(lines + 1); out->score.reserve(lines + 1); out->strand.reserve(lines + 1);
        out->start.reserve(lines + 1); out->end.reserve(lines + 1);
    }
    bool header = true;
    size_t n_fields0 = 0;
    Tok fld[16];
    for (size_t pos = 0; pos < total;) {
        char* line = &b[pos];
        char* nl = static_cast<char*>(memchr(line, '\n', b.size() - pos));
        size_t len = (size_t)(nl - line);
        pos += len + 1;
        if (header && is_header(line, len)) continue;                         // GetHeader
        // GetHeader reads the first data line itself; if that getline hit the end of the file (no trailing newline) the
        // stream is no longer good() and GetNextBed returns at once: a lone unterminated data line is never seen
        if (header && pos > total) break;
        header = false;
        if (len && line[len - 1] == '\r') --len;
        line[len] = '\0';
        size_t nf = 0;                                                        // number of fields (all counted, 16 kept)
        if (len) {
            char* q = line;
            for (;;) {
                char* t = static_cast<char*>(memchr(q, '\t', (size_t)(line + len - q)));
                char* e = t ? t : line + len;
                if (!t && e == q) break;                                      // the line ended with a tab: no trailing empty field
                if (nf < 16) fld[nf] = Tok{q, (uint32_t)(e - q)};
                ++nf;
                *e = '\0';
                if (!t) break;
                q = t + 1;
            }
        }
        if (nf == 0 || is_header(fld[0].p, fld[0].n)) break;                  // BED_BLANK / BED_HEADER end get_single_junction's loop
        if (nf < 3) { out->stop_is_exit = true; out->stop = fail(RTJX_E_IO, "It looks as though you have less than 3 columns. Are you sure your files are tab-delimited?"); break; }
        if (!is_integer(fld[1].p) || !is_integer(fld[2].p)) { out->stop_is_exit = true; out->stop = fail(RTJX_E_IO, "Unexpected file format.  Please use tab-delimited BED, GFF, or VCF."); break; }
        if (!n_fields0) n_fields0 = nf;
        if (nf != n_fields0) { out->stop_is_exit = true; out->stop = fail(RTJX_E_IO, "Differing number of BED fields encountered. Exiting..."); break; }
        // parseBedLine (bedFile.h:685-760): negative coordinates and start > end end the run (exit 1), start == end is widened
        const int i_start = atoi(fld[1].p), i_end = atoi(fld[2].p);
        if (i_start < 0) { out->stop_is_exit = true; out->stop = fail(RTJX_E_IO, "Error: malformed BED entry. Start Coordinate detected that is < 0. Exiting."); break; }
        if (i_end < 0) { out->stop_is_exit = true; out->stop = fail(RTJX_E_IO, "Error: malformed BED entry. End Coordinate detected that is < 0. Exiting."); break; }
        uint32_t start = (uint32_t)i_start, end = (uint32_t)i_end;
        if (start == end) { --start; ++end; }                                 // zeroLength
        if (start > end) { out->stop_is_exit = true; out->stop = fail(RTJX_E_IO, "Error: malformed BED entry. Start was greater than end. Exiting."); break; }
        if (nf != 12 || fld[10].n == 0) {                                     // :70-75
            out->stop = fail(RTJX_E_IO, "BED line not in BED12 format. start: " + std::string(fld[0].p, fld[0].n) + ":" + std::to_string(start));
            break;
        }
        const char* comma = static_cast<const char*>(memchr(fld[10].p, ',', fld[10].n));      // Tokenize(blocksize_field, ints, ',')
        start += (uint32_t)atoi(fld[10].p);                                   // atoi stops at the comma
        end -= (uint32_t)(((comma && comma[1]) ? atoi(comma + 1) : 0) - 1);
        out->chrom.push_back(fld[0]); out->name.push_back(fld[3]); out->score.push_back(fld[4]); out->strand.push_back(fld[5]);
        out->start.push_back(start); out->end.push_back(end);
    }
    return Status();
}

// ---- device buffers ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> cudaError_t upload(const std::vector<T>& v, cudaStream_t st) {
        const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return e;
        if (!v.empty()) e = cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
        return e;
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

bool write_all(int fd, const std::string& s) {
    size_t off = 0;
    while (off < s.size()) {
        const ssize_t w = ::write(fd, s.data() + off, s.size() - off);
        if (w <= 0) return false;
        off += (size_t)w;
    }
    return true;
}

#define ACK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return fail(RTJX_E_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e__)); \
    } while (0)

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

Status annotate_impl(const rtjx_annotate_params& p, int out_fd, uint64_t* n_lines) {
    if (!p.junctions_bed || !p.fasta || !p.gtf) return fail(RTJX_E_ARG, "Error parsing inputs!(2)\n\n");
    const bool trace = getenv("RTJX_TRACE") != nullptr;                       // stage timings on stderr
    double t_prev = now_s();
    auto lap = [&](const char* what) { if (trace) { const double t = now_s(); fprintf(stderr, "[rtjx annotate] %-22s %8.2f ms\n", what, (t - t_prev) * 1e3); t_prev = t; } };
    // junctions_main.cc:68-72: options, GTF, junctions file, then the output stream and the header
    GtfFlat gtf;
    Status st = load_gtf(p.gtf, &gtf);
    if (!st.ok()) return st;
    lap("gtf load");
    JunctionLines jl;
    if (!(st = read_junctions(p.junctions_bed, &jl)).ok()) return st;
    const size_t n = jl.start.size();
    lap("junctions read");

    // the annotation itself runs on the device only
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev == 0)
        return fail(RTJX_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(ce) +
                                     "); regtools-b200 has no CPU fallback for junctions annotate");
    if (p.device < 0 || p.device >= n_dev) return fail(RTJX_E_CUDA, "CUDA device ordinal out of range");
    ACK(cudaSetDevice(p.device));

    struct FdGuard { int fd = -1; ~FdGuard() { if (fd >= 0) ::close(fd); } } own;
    if (out_fd < 0) {                                                         // set_ofstream_object (:41-51)
        if (!p.out_path) return fail(RTJX_E_ARG, "no output: out_fd < 0 and out_path is NULL");
        own.fd = ::open(p.out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (own.fd < 0) return fail(RTJX_E_IO, std::string("Unable to open ") + p.out_path);
        out_fd = own.fd;
    }
    static const char HEADER[] = "chrom\tstart\tend\tname\tscore\tstrand\tsplice_site\tacceptors_skipped\texons_skipped\tdonors_skipped\tanchor"
                                 "\tknown_donor\tknown_acceptor\tknown_junction\tgene_names\tgene_ids\ttranscripts\n";
    std::vector<AnnOut> res(n);
    std::vector<unsigned long long> items;
    if (n) {
        // chrom dictionary of the junctions -> GTF chrom id and FASTA sequence
        FastaGenome genome;
        std::string ferr;
        const bool have_fasta = load_fasta(p.fasta, &genome, &ferr);
        lap("fasta load");
        std::map<std::string, int32_t> cdict;
        std::vector<int32_t> c_gtf, j_chrom(n);
        std::vector<unsigned long long> c_goff, c_glen;
        std::vector<uint8_t> j_strand(n);
        Tok last{nullptr, 0}; int32_t last_id = -1;                             // a BED is usually sorted: one lookup per run of a chrom
        for (size_t i = 0; i < n; ++i) {
            const Tok c = jl.chrom[i];
            if (!(last.p && last.n == c.n && memcmp(last.p, c.p, c.n) == 0)) {
                const std::string name(c.p, c.n);
                std::map<std::string, int32_t>::iterator it = cdict.find(name);
                if (it == cdict.end()) {
                    it = cdict.insert(std::make_pair(name, (int32_t)cdict.size())).first;
                    std::map<std::string, uint32_t>::const_iterator g = gtf.chrom_id.find(name);
                    c_gtf.push_back(g == gtf.chrom_id.end() ? -1 : (int32_t)g->second);
                    const int q = have_fasta ? genome.find(name) : -1;
                    c_goff.push_back(q >= 0 ? genome.offset[(size_t)q] : 0ull);
                    c_glen.push_back(q >= 0 ? genome.length[(size_t)q] : ~0ull);
                }
                last = c; last_id = it->second;
            }
            j_chrom[i] = last_id;
            const Tok st = jl.strand[i];
            j_strand[i] = (st.n == 1 && st.p[0] == '+') ? 0 : ((st.n == 1 && st.p[0] == '-') ? 1 : 2);
        }
        cudaStream_t stream = nullptr;
        ACK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{stream};
        DevBuf d_bin_key, d_bin_off, d_bin_tx, d_tx_off, d_tx_strand, d_ex_start, d_ex_end, d_js, d_je, d_jstrand, d_jchrom, d_cgtf, d_cgoff,
            d_cglen, d_genome, d_out, d_items, d_ctr;
        ACK(d_bin_key.upload(gtf.bin_key, stream)); ACK(d_bin_off.upload(gtf.bin_off, stream)); ACK(d_bin_tx.upload(gtf.bin_tx, stream));
        ACK(d_tx_off.upload(gtf.tx_ex_off, stream)); ACK(d_tx_strand.upload(gtf.tx_strand, stream));
        ACK(d_ex_start.upload(gtf.ex_start, stream)); ACK(d_ex_end.upload(gtf.ex_end, stream));
        ACK(d_js.upload(jl.start, stream)); ACK(d_je.upload(jl.end, stream)); ACK(d_jstrand.upload(j_strand, stream));
        ACK(d_jchrom.upload(j_chrom, stream)); ACK(d_cgtf.upload(c_gtf, stream)); ACK(d_cgoff.upload(c_goff, stream));
        ACK(d_cglen.upload(c_glen, stream)); ACK(d_genome.upload(genome.bases, stream));
        ACK(d_out.alloc(n * sizeof(AnnOut))); ACK(d_ctr.alloc(ANN_CTR_COUNT * sizeof(uint32_t)));

        AnnGtfView gv;
        gv.bin_key = d_bin_key.as<unsigned long long>(); gv.bin_off = d_bin_off.as<uint32_t>(); gv.bin_tx = d_bin_tx.as<uint32_t>();
        gv.n_bins = (uint32_t)gtf.bin_key.size();
        gv.tx_ex_off = d_tx_off.as<uint32_t>(); gv.tx_strand = d_tx_strand.as<uint8_t>();
        gv.ex_start = d_ex_start.as<uint32_t>(); gv.ex_end = d_ex_end.as<uint32_t>(); gv.n_tx = (uint32_t)gtf.tx_id.size();
        AnnJunctionView jv;
        jv.start = d_js.as<uint32_t>(); jv.end = d_je.as<uint32_t>(); jv.strand = d_jstrand.as<uint8_t>(); jv.chrom = d_jchrom.as<int32_t>();
        jv.n = (uint32_t)n;
        jv.c_gtf = d_cgtf.as<int32_t>(); jv.c_goff = d_cgoff.as<unsigned long long>(); jv.c_glen = d_cglen.as<unsigned long long>();
        jv.genome = d_genome.as<uint8_t>();

        lap("upload");
        unsigned long long cap = 16ull * n + (1ull << 16);
        if (const char* v = getenv("RTJX_ANNOTATE_ITEMS")) cap = strtoull(v, nullptr, 10);     // tests: force the grow-and-rerun path
        uint32_t ctr[ANN_CTR_COUNT];
        for (int attempt = 0;; ++attempt) {
            if (d_items.p) { cudaFree(d_items.p); d_items.p = nullptr; }
            ACK(d_items.alloc((size_t)cap * sizeof(unsigned long long)));
            const uint32_t init[ANN_CTR_COUNT] = {0u, 0u, 0u, 0xffffffffu};
            ACK(cudaMemcpyAsync(d_ctr.p, init, sizeof init, cudaMemcpyHostToDevice, stream));
            launch_annotate(gv, jv, p.include_single_exon ? 0 : 1, d_items.as<unsigned long long>(), cap, d_out.as<AnnOut>(), d_ctr.as<uint32_t>(), stream);
            ACK(cudaGetLastError());
            ACK(cudaMemcpyAsync(ctr, d_ctr.p, sizeof ctr, cudaMemcpyDeviceToHost, stream));
            ACK(cudaStreamSynchronize(stream));
            if (!ctr[ANN_CTR_OVERFLOW]) break;
            if (attempt) return fail(RTJX_E_STATE, "internal: annotation item buffer overflowed twice");
            cap = ((unsigned long long)ctr[ANN_CTR_CURSOR + 1] << 32 | ctr[ANN_CTR_CURSOR]) + 16;   // the exact need
        }
        const unsigned long long used = (unsigned long long)ctr[ANN_CTR_CURSOR + 1] << 32 | ctr[ANN_CTR_CURSOR];
        items.resize((size_t)used);
        ACK(cudaMemcpyAsync(res.data(), d_out.p, n * sizeof(AnnOut), cudaMemcpyDeviceToHost, stream));
        if (used) ACK(cudaMemcpyAsync(items.data(), d_items.p, (size_t)used * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        ACK(cudaStreamSynchronize(stream));
        lap("kernel + download");
    }

    // ---- AnnotatedJunction::print (junctions_annotator.h:86-121), in file order; errors surface where the reference throws
    // genes_overlap is a set<vector<string>>: (gene_name, gene_id) pairs in lexicographic order -> rank every transcript's pair once
    std::vector<std::pair<std::string, std::string> > gene_pairs;
    std::vector<uint32_t> gene_rank(gtf.tx_id.size());
    {
        std::map<std::pair<std::string, std::string>, uint32_t> rank;
        for (size_t t = 0; t < gtf.tx_id.size(); ++t) rank[std::make_pair(gtf.gene_name[t], gtf.gene_id[t])] = 0;
        uint32_t r = 0;
        for (std::map<std::pair<std::string, std::string>, uint32_t>::iterator it = rank.begin(); it != rank.end(); ++it) { it->second = r++; gene_pairs.push_back(it->first); }
        for (size_t t = 0; t < gtf.tx_id.size(); ++t) gene_rank[t] = rank[std::make_pair(gtf.gene_name[t], gtf.gene_id[t])];
    }
    struct TextBuf {
        std::string s;
        void u32(uint32_t v) { char t[10]; int k = 0; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); while (k) s.push_back(t[--k]); }
        void tok(const Tok& t) { s.append(t.p, t.n); }
        void position(const Tok& chrom, uint32_t a, uint32_t b) { tok(chrom); s.push_back(':'); u32(a); s.push_back('-'); u32(b); }
    } out, chatter;
    out.s.reserve((1u << 20) + 4096);
    out.s += HEADER;
    const bool chat = p.chatter_fd >= 0;
    uint64_t printed = 0;
    Status result;
    std::vector<uint32_t> ranks;
    for (size_t i = 0; i < n; ++i) {
        const AnnOut& r = res[i];
        if (chat) { chatter.s += "position = "; chatter.position(jl.chrom[i], jl.start[i] + 1u, jl.start[i] + 2u); chatter.s.push_back('\n'); }
        if (r.flags & ANN_NO_CONTIG) {                                         // get_reference_sequence :391-402
            TextBuf p1;
            p1.position(jl.chrom[i], jl.start[i] + 1u, jl.start[i] + 2u);
            result = fail(RTJX_E_IO, "Unable to extract FASTA sequence for position " + p1.s + "\n\n");
            break;
        }
        if (chat) { chatter.s += "position = "; chatter.position(jl.chrom[i], jl.end[i] - 2u, jl.end[i] - 1u); chatter.s.push_back('\n'); }
        std::string& o = out.s;
        out.tok(jl.chrom[i]); o.push_back('\t'); out.u32(jl.start[i]); o.push_back('\t'); out.u32(jl.end[i]); o.push_back('\t');
        out.tok(jl.name[i]); o.push_back('\t'); out.tok(jl.score[i]); o.push_back('\t'); out.tok(jl.strand[i]); o.push_back('\t');
        o.append(reinterpret_cast<const char*>(r.ss), r.ss_n[0]); o.push_back('-'); o.append(reinterpret_cast<const char*>(r.ss + 3), r.ss_n[1]);
        o.push_back('\t'); out.u32(r.n_acceptors); o.push_back('\t'); out.u32(r.n_exons); o.push_back('\t'); out.u32(r.n_donors);
        const bool kd = r.flags & ANN_KNOWN_DONOR, ka = r.flags & ANN_KNOWN_ACCEPTOR, kj = r.flags & ANN_KNOWN_JUNCTION;
        o.push_back('\t'); o += kj ? "DA" : (kd ? (ka ? "NDA" : "D") : (ka ? "A" : "N"));               // annotate_anchor :314-328
        o.push_back('\t'); o.push_back(kd ? '1' : '0'); o.push_back('\t'); o.push_back(ka ? '1' : '0'); o.push_back('\t'); o.push_back(kj ? '1' : '0');
        if (r.n_tx) {
            const unsigned long long off = (unsigned long long)r.tx_off_hi << 32 | r.tx_off_lo;
            ranks.clear();
            for (uint32_t k = 0; k < r.n_tx; ++k) ranks.push_back(gene_rank[(size_t)items[(size_t)off + k]]);
            std::sort(ranks.begin(), ranks.end());
            ranks.erase(std::unique(ranks.begin(), ranks.end()), ranks.end());
            o.push_back('\t');
            for (size_t k = 0; k < ranks.size(); ++k) { if (k) o.push_back(','); o += gene_pairs[ranks[k]].first; }
            o.push_back('\t');
            for (size_t k = 0; k < ranks.size(); ++k) { if (k) o.push_back(','); o += gene_pairs[ranks[k]].second; }
            o.push_back('\t');
            for (uint32_t k = 0; k < r.n_tx; ++k) { if (k) o.push_back(','); o += gtf.tx_id[(size_t)items[(size_t)off + k]]; }
        } else {
            o += "\tNA\tNA\tNA";
        }
        o.push_back('\n');
        ++printed;
        if (o.size() > (1u << 20)) { if (!write_all(out_fd, o)) return fail(RTJX_E_IO, "write failed"); o.clear(); }
        if (chatter.s.size() > (1u << 20)) { write_all(p.chatter_fd, chatter.s); chatter.s.clear(); }
    }
    const bool header_lost = own.fd >= 0 && printed == 0 && result.ok() && jl.stop_is_exit;      // see JunctionLines::stop_is_exit
    if (!header_lost && !write_all(out_fd, out.s)) return fail(RTJX_E_IO, "write failed");
    lap("format + write");
    if (result.ok() && !jl.stop.ok()) result = jl.stop;                        // a malformed BED line ends the run after the lines before it
    if (chat) {
        if (result.ok()) chatter.s += "\nAnnotated " + std::to_string(printed) + " lines.\n";
        write_all(p.chatter_fd, chatter.s);
    }
    if (n_lines) *n_lines = printed;
    return result;
}

}  // namespace
}  // namespace rtjx

extern "C" {

void rtjx_annotate_params_default(rtjx_annotate_params* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->struct_size = (uint32_t)sizeof *p;
    p->chatter_fd = -1;
}

int rtjx_annotate(const rtjx_annotate_params* p, int out_fd, uint64_t* n_lines, char* err, size_t err_cap) {
    if (err && err_cap) err[0] = '\0';
    if (!p || p->struct_size != sizeof(rtjx_annotate_params)) return RTJX_E_ARG;
    rtjx::Status st;
    try { st = rtjx::annotate_impl(*p, out_fd, n_lines); }
    catch (const std::bad_alloc&) { st = rtjx::fail(RTJX_E_NOMEM, "out of host memory"); }
    catch (const std::exception& e) { st = rtjx::fail(RTJX_E_STATE, e.what()); }
    catch (...) { st = rtjx::fail(RTJX_E_STATE, "unknown internal error"); }
    if (!st.ok() && err && err_cap) { strncpy(err, st.msg.c_str(), err_cap - 1); err[err_cap - 1] = '\0'; }
    return st.code;
}

}  // extern "C"
