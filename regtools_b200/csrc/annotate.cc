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
struct JunctionLines {
    std::vector<std::string> chrom, name, score, strand;
    std::vector<uint32_t> start, end;                                         // adjusted
    Status stop;                                                              // why reading stopped early (error after the lines above)
};

bool is_integer(const std::string& s) {
    if (s.empty()) return false;
    char* e = nullptr;
    strtol(s.c_str(), &e, 10);
    return *e == '\0';
}
bool is_header(const std::string& s) { return s.find("#") == 0 || s.find("browser") == 0 || s.find("track") == 0; }

Status read_junctions(const std::string& path, JunctionLines* out) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) return fail(RTJX_E_IO, "Error: The requested file (" + path + ") could not be opened. Exiting!");
    std::string line;
    bool header = true;
    size_t n_fields0 = 0;
    while (std::getline(f, line)) {
        if (header && is_header(line)) continue;                              // GetHeader
        header = false;
        if (!line.empty() && line[line.size() - 1] == '\r') line.resize(line.size() - 1);
        const std::vector<std::string> fld = split(line, '\t');
        if (fld.empty() || is_header(fld[0])) break;                          // BED_BLANK / BED_HEADER end get_single_junction's loop
        if (fld.size() < 3) { out->stop = fail(RTJX_E_IO, "It looks as though you have less than 3 columns. Are you sure your files are tab-delimited?"); break; }
        if (!is_integer(fld[1]) || !is_integer(fld[2])) { out->stop = fail(RTJX_E_IO, "Unexpected file format.  Please use tab-delimited BED, GFF, or VCF."); break; }
        if (!n_fields0) n_fields0 = fld.size();
        if (fld.size() != n_fields0) { out->stop = fail(RTJX_E_IO, "Differing number of BED fields encountered. Exiting..."); break; }
        uint32_t start = (uint32_t)atoi(fld[1].c_str()), end = (uint32_t)atoi(fld[2].c_str());
        if (fld.size() != 12 || fld[10].empty()) {                            // :70-75
            out->stop = fail(RTJX_E_IO, "BED line not in BED12 format. start: " + fld[0] + ":" + std::to_string(start));
            break;
        }
        const std::vector<std::string> bs = split(fld[10], ',');
        start += (uint32_t)atoi(bs[0].c_str());
        end -= (uint32_t)(atoi(bs.size() > 1 ? bs[1].c_str() : "0") - 1);
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

Status annotate_impl(const rtjx_annotate_params& p, int out_fd, uint64_t* n_lines) {
    if (!p.junctions_bed || !p.fasta || !p.gtf) return fail(RTJX_E_ARG, "Error parsing inputs!(2)\n\n");
    // junctions_main.cc:68-72: options, GTF, junctions file, then the output stream and the header
    GtfFlat gtf;
    Status st = load_gtf(p.gtf, &gtf);
    if (!st.ok()) return st;
    JunctionLines jl;
    if (!(st = read_junctions(p.junctions_bed, &jl)).ok()) return st;
    const size_t n = jl.start.size();

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
        std::map<std::string, int32_t> cdict;
        std::vector<int32_t> c_gtf, j_chrom(n);
        std::vector<unsigned long long> c_goff, c_glen;
        std::vector<uint8_t> j_strand(n);
        for (size_t i = 0; i < n; ++i) {
            std::map<std::string, int32_t>::iterator it = cdict.find(jl.chrom[i]);
            if (it == cdict.end()) {
                it = cdict.insert(std::make_pair(jl.chrom[i], (int32_t)cdict.size())).first;
                std::map<std::string, uint32_t>::const_iterator g = gtf.chrom_id.find(jl.chrom[i]);
                c_gtf.push_back(g == gtf.chrom_id.end() ? -1 : (int32_t)g->second);
                const int q = have_fasta ? genome.find(jl.chrom[i]) : -1;
                c_goff.push_back(q >= 0 ? genome.offset[(size_t)q] : 0ull);
                c_glen.push_back(q >= 0 ? genome.length[(size_t)q] : ~0ull);
            }
            j_chrom[i] = it->second;
            j_strand[i] = jl.strand[i] == "+" ? 0 : (jl.strand[i] == "-" ? 1 : 2);
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
    }

    // ---- AnnotatedJunction::print (junctions_annotator.h:86-121), in file order; errors surface where the reference throws
    std::string out;
    out.reserve(1u << 20);
    out += HEADER;
    std::string chatter;
    uint64_t printed = 0;
    Status result;
    for (size_t i = 0; i < n; ++i) {
        const AnnOut& r = res[i];
        const std::string p1 = jl.chrom[i] + ":" + std::to_string(jl.start[i] + 1u) + "-" + std::to_string(jl.start[i] + 2u);
        if (p.chatter_fd >= 0) chatter += "position = " + p1 + "\n";
        if (r.flags & ANN_NO_CONTIG) {                                         // get_reference_sequence :391-402
            result = fail(RTJX_E_IO, "Unable to extract FASTA sequence for position " + p1 + "\n\n");
            break;
        }
        if (p.chatter_fd >= 0)
            chatter += "position = " + jl.chrom[i] + ":" + std::to_string(jl.end[i] - 2u) + "-" + std::to_string(jl.end[i] - 1u) + "\n";
        out += jl.chrom[i]; out += '\t'; out += std::to_string(jl.start[i]); out += '\t'; out += std::to_string(jl.end[i]); out += '\t';
        out += jl.name[i]; out += '\t'; out += jl.score[i]; out += '\t'; out += jl.strand[i]; out += '\t';
        out.append(reinterpret_cast<const char*>(r.ss), r.ss_n[0]); out += '-'; out.append(reinterpret_cast<const char*>(r.ss + 3), r.ss_n[1]);
        out += '\t'; out += std::to_string(r.n_acceptors); out += '\t'; out += std::to_string(r.n_exons); out += '\t'; out += std::to_string(r.n_donors);
        const bool kd = r.flags & ANN_KNOWN_DONOR, ka = r.flags & ANN_KNOWN_ACCEPTOR, kj = r.flags & ANN_KNOWN_JUNCTION;
        out += '\t'; out += kj ? "DA" : (kd ? (ka ? "NDA" : "D") : (ka ? "A" : "N"));                  // annotate_anchor :314-328
        out += '\t'; out += kd ? '1' : '0'; out += '\t'; out += ka ? '1' : '0'; out += '\t'; out += kj ? '1' : '0';
        if (r.n_tx) {
            const unsigned long long off = (unsigned long long)r.tx_off_hi << 32 | r.tx_off_lo;
            std::set<std::pair<std::string, std::string> > genes;              // set<vector<string>>: (gene_name, gene_id) lexicographic
            for (uint32_t k = 0; k < r.n_tx; ++k) {
                const size_t t = (size_t)items[(size_t)off + k];
                genes.insert(std::make_pair(gtf.gene_name[t], gtf.gene_id[t]));
            }
            out += '\t';
            for (std::set<std::pair<std::string, std::string> >::iterator it = genes.begin(); it != genes.end(); ++it) { if (it != genes.begin()) out += ','; out += it->first; }
            out += '\t';
            for (std::set<std::pair<std::string, std::string> >::iterator it = genes.begin(); it != genes.end(); ++it) { if (it != genes.begin()) out += ','; out += it->second; }
            out += '\t';
            for (uint32_t k = 0; k < r.n_tx; ++k) { if (k) out += ','; out += gtf.tx_id[(size_t)items[(size_t)off + k]]; }
        } else {
            out += "\tNA\tNA\tNA";
        }
        out += '\n';
        ++printed;
        if (out.size() > (1u << 20)) { if (!write_all(out_fd, out)) return fail(RTJX_E_IO, "write failed"); out.clear(); }
    }
    if (!write_all(out_fd, out)) return fail(RTJX_E_IO, "write failed");
    if (result.ok() && !jl.stop.ok()) result = jl.stop;                        // a malformed BED line ends the run after the lines before it
    if (p.chatter_fd >= 0) {
        if (result.ok()) chatter += "\nAnnotated " + std::to_string(printed) + " lines.\n";
        write_all(p.chatter_fd, chatter);
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
