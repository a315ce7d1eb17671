// pcramp_b200 -- the command-line host of the B200 path: pcramp's design loop (main.cpp:56-1300) with every data-parallel step
// behind the C ABI of include/pcramp_gpu.h (libpcramp_gpu.so).  Plain C++ (g++), no CUDA in this file.
//
// Kept from the reference: the option names of its command line (options.cpp:150-215; the subset below), FASTA input (plain or
// gzip through zlib, as parse_fasta.cpp reads it), the text report (main.cpp:131-147, 520-523, 946-1075, 1131-1192) -- so that
// the file written for `--seed S --thread 1` can be compared line by line with the stock program's -- and the progress lines on
// stderr.  Not offered (refused with a message): -T / -B directory groups, --json input, --o.json output, --optimize.top-down.
//
//   pcramp_b200 -t targets.fa [-b background.fa] -o out.txt [--seed S] [--trial N] [--count K] [-d D] [--optimize.5] [--optimize.3] ...
//
// --thread k selects the seed streams of candidate generation (k = 1 reproduces the stock program at --thread 1 exactly;
// larger k = the static schedule of k OpenMP threads, their seeds drawn in thread order).  --device picks the GPU.
#include "../../include/pcramp_gpu.h"

#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct HostOptions {
	pcramp_gpu_design_options d;
	std::vector<std::string> target_files, background_files, target_ignore, background_ignore;
	std::string output;
	unsigned int seed = 0, num_assay = 100; // DEFAULT_NUM_ASSAY (pcramp.h:33)
	int target_len_min = 0, target_len_max = 2147483647, background_len_min = 0, background_len_max = 2147483647;
	bool normalize_target = false, normalize_background = false, silent = false, timing = false;
	int device = 0;
};

[[noreturn]] void die(const std::string &m)
{
	std::cerr << "pcramp_b200: " << m << std::endl;
	exit(EXIT_FAILURE);
}

std::string read_inflated(const std::string &path)
{ // gzopen reads plain files too (parse_fasta.cpp:20-27 opens every input that way)
	gzFile fin = gzopen(path.c_str(), "r");
	if (!fin) die("Unable to open " + path);
	std::string out;
	std::vector<char> buf(1 << 20);
	int n;
	while ((n = gzread(fin, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), (size_t)n);
	gzclose(fin);
	return out;
}

std::string word_string(const uint64_t w[2], bool lower)
{ // operator<<(ostream, Word) (word.h:649-657) / tolower(Word::str())
	char s[33];
	pcramp_word_to_string(w, s);
	std::string out(s);
	if (lower)
		for (char &c : out) c = (char)tolower(c);
	return out;
}

struct Loaded {
	uint32_t n = 0;
	std::vector<std::string> defline;
	std::vector<uint32_t> length;
};

// parse_fasta over the files (main.cpp:257-279 / :348-370) with the optional per-file weight normalisation
Loaded load_collection(pcramp_gpu_ctx *ctx, int kind, const std::vector<std::string> &files, int min_len, int max_len,
	const std::vector<std::string> &ignore, bool normalize)
{
	Loaded L;
	std::vector<std::string> text(files.size());
	std::vector<const char *> ptr(files.size()), ig(ignore.size());
	std::vector<uint64_t> bytes(files.size());
	for (size_t i = 0; i < files.size(); ++i) {
		text[i] = read_inflated(files[i]);
		ptr[i] = text[i].data();
		bytes[i] = text[i].size();
	}
	for (size_t i = 0; i < ignore.size(); ++i) ig[i] = ignore[i].c_str();
	uint32_t n = 0;
	if (pcramp_gpu_upload_fasta(ctx, kind, (uint32_t)files.size(), ptr.data(), bytes.data(), (uint64_t)std::max(0, min_len), (uint64_t)max_len,
			(uint32_t)ignore.size(), ig.data(), &n))
		die(pcramp_gpu_last_error(ctx));
	L.n = n;
	std::vector<uint32_t> file(n), dlen(n);
	std::vector<uint64_t> doff(n);
	std::vector<float> weight(n);
	L.length.resize(n);
	if (n && pcramp_gpu_fasta_records(ctx, kind, file.data(), doff.data(), dlen.data(), L.length.data(), weight.data())) die(pcramp_gpu_last_error(ctx));
	for (uint32_t i = 0; i < n; ++i) L.defline.push_back(text[file[i]].substr(doff[i], dlen[i]));
	if (normalize && n) { // main.cpp:268-278
		std::vector<uint32_t> per_file(files.size(), 0);
		for (uint32_t i = 0; i < n; ++i) per_file[file[i]]++;
		for (uint32_t i = 0; i < n; ++i) weight[i] = (float)(1.0 / per_file[file[i]]);
		if (pcramp_gpu_set_weights(ctx, kind, weight.data())) die(pcramp_gpu_last_error(ctx));
	}
	pcramp_gpu_fasta_free(ctx);
	return L;
}

void sequence_summary(const std::string &prefix, const Loaded &L, std::ostream &out)
{ // main.cpp:1305-1399, text format
	out << prefix << " Number of sequences = " << L.n << std::endl;
	if (!L.n) return;
	float ave = 0.0f, stdev = 0.0f;
	unsigned int lo = L.length[0], hi = L.length[0];
	for (uint32_t l : L.length) {
		ave += l;
		lo = std::min(lo, l);
		hi = std::max(hi, l);
	}
	ave /= L.n;
	for (uint32_t l : L.length) {
		const float tmp = l - ave;
		stdev += tmp * tmp;
	}
	stdev = (L.n > 1) ? std::sqrt(stdev / (L.n - 1)) : 0.0f;
	out << prefix << " Min sequence length = " << lo << std::endl;
	out << prefix << " Max sequence length = " << hi << std::endl;
	out << prefix << " Average sequence length = " << ave << std::endl;
	out << prefix << " Stdev sequence length = " << stdev << std::endl;
}

void usage()
{
	std::cerr << "pcramp_b200 (PCRamp 0.3 design loop on the B200 path)\n"
	             "\t-t <target fasta file> (repeatable)\n\t[-b <background fasta file>] (repeatable)\n\t-o <output file>\n"
	             "\t[-d <max degeneracy>] [--trial <n>] [--seed <s>] [--thread <seed streams>] [--count <assays>]\n"
	             "\t[--optimize.5] [--no-optimize.5] [--optimize.3] [--no-optimize.3] [--salt <M>]\n"
	             "\t[--primer.hairpin <Tm>] [--primer.dimer <Tm>] [--primer.size.min|max <n>] [--primer.tm.min|max <Tm>] [--primer.strand <M>]\n"
	             "\t[--primer.taq-mama] [--target.amplicon.min|max <n>] [--target.threshold <x>] [--target.search <x>] [--target.cover <x>]\n"
	             "\t[--target.ignore <key>] [--target.normalize] [--target.size.min|max <n>]\n"
	             "\t[--background.amplicon.min|max <n>] [--background.threshold <x>] [--background.search <x>] [--background.cover <x>]\n"
	             "\t[--background.ignore <key>] [--background.normalize] [--background.size.min|max <n>]\n"
	             "\t[--pack.degen.max <n>] [--pack.gc.min|max <x>] [-v silent|verbose] [--device <gpu>] [--timing]\n";
}

HostOptions parse(int argc, char **argv)
{
	HostOptions h;
	pcramp_gpu_design_default_options(&h.d);
	auto need = [&](int &i) -> const char * {
		if (i + 1 >= argc) die(std::string("missing value for ") + argv[i]);
		return argv[++i];
	};
	for (int i = 1; i < argc; ++i) {
		const std::string a = argv[i];
		if (a == "-t") h.target_files.push_back(need(i));
		else if (a == "-b") h.background_files.push_back(need(i));
		else if (a == "-o") h.output = need(i);
		else if (a == "-d") h.d.degen = (uint32_t)atoi(need(i));
		else if (a == "-v") h.silent = std::string(need(i)) == "silent";
		else if (a == "--trial") h.d.num_trial = (uint32_t)atoi(need(i));
		else if (a == "--seed") h.seed = (unsigned int)strtoul(need(i), nullptr, 10);
		else if (a == "--thread") h.d.n_streams = (uint32_t)std::max(1, atoi(need(i)));
		else if (a == "--count") h.num_assay = (unsigned int)atoi(need(i));
		else if (a == "--optimize.5") h.d.optimize_5 = 1;
		else if (a == "--no-optimize.5") h.d.optimize_5 = 0;
		else if (a == "--optimize.3") h.d.optimize_3 = 1;
		else if (a == "--no-optimize.3") h.d.optimize_3 = 0;
		else if (a == "--salt") h.d.salt = (float)atof(need(i));
		else if (a == "--primer.hairpin") h.d.max_hairpin = (float)atof(need(i));
		else if (a == "--primer.dimer") h.d.max_dimer = (float)atof(need(i));
		else if (a == "--primer.size.min") h.d.primer_min = atoi(need(i));
		else if (a == "--primer.size.max") h.d.primer_max = atoi(need(i));
		else if (a == "--primer.tm.min") h.d.primer_tm_min = (float)atof(need(i));
		else if (a == "--primer.tm.max") h.d.primer_tm_max = (float)atof(need(i));
		else if (a == "--primer.strand") h.d.primer_strand = (float)atof(need(i));
		else if (a == "--primer.taq-mama") h.d.use_taq_mama = 1;
		else if (a == "--target.amplicon.min") h.d.target_amplicon_min = atoi(need(i));
		else if (a == "--target.amplicon.max") h.d.target_amplicon_max = atoi(need(i));
		else if (a == "--target.threshold") h.d.target_threshold = (float)atof(need(i));
		else if (a == "--target.search") h.d.target_search_multiplier = (float)atof(need(i));
		else if (a == "--target.cover") h.d.min_target_cover = (float)atof(need(i));
		else if (a == "--target.ignore") h.target_ignore.push_back(need(i));
		else if (a == "--target.normalize") h.normalize_target = true;
		else if (a == "--target.size.min") h.target_len_min = atoi(need(i));
		else if (a == "--target.size.max") h.target_len_max = atoi(need(i));
		else if (a == "--background.amplicon.min") h.d.background_amplicon_min = atoi(need(i));
		else if (a == "--background.amplicon.max") h.d.background_amplicon_max = atoi(need(i));
		else if (a == "--background.threshold") h.d.background_threshold = (float)atof(need(i));
		else if (a == "--background.search") h.d.background_search_multiplier = (float)atof(need(i));
		else if (a == "--background.cover") h.d.max_background_cover = (float)atof(need(i));
		else if (a == "--background.ignore") h.background_ignore.push_back(need(i));
		else if (a == "--background.normalize") h.normalize_background = true;
		else if (a == "--background.size.min") h.background_len_min = atoi(need(i));
		else if (a == "--background.size.max") h.background_len_max = atoi(need(i));
		else if (a == "--pack.degen.max") h.d.pack_max_degen = (uint32_t)atoi(need(i));
		else if (a == "--pack.gc.min") h.d.pack_min_gc = (float)atof(need(i));
		else if (a == "--pack.gc.max") h.d.pack_max_gc = (float)atof(need(i));
		else if (a == "--o.text") {}
		else if (a == "--device") h.device = atoi(need(i));
		else if (a == "--timing") h.timing = true;
		else if (a == "-h" || a == "--help" || a == "-?") { usage(); exit(EXIT_SUCCESS); }
		else if (a == "-T" || a == "-B" || a == "--json" || a == "--json.root" || a == "--o.json" || a == "--optimize.top-down" ||
		         a.find(".prefix") != std::string::npos)
			die(a + " is not offered by this host (directory groups, JSON input / output and the top-down search stay with the stock program)");
		else die("unknown option " + a);
	}
	if (argc <= 1) { usage(); exit(EXIT_SUCCESS); }
	if (h.target_files.empty()) die("Please specify one or more target sequences (-t)");
	if (h.output.empty()) die("Please specify an output file (-o)");
	if (h.seed == 0) h.seed = (unsigned int)time(nullptr); // options.cpp: a seed of 0 triggers a time-based seed
	return h;
}

} // namespace

int main(int argc, char **argv)
{
	const auto t0 = std::chrono::steady_clock::now();
	HostOptions h = parse(argc, argv);
	std::ofstream fnull("/dev/null");
	std::ostream &vout = h.silent ? (std::ostream &)fnull : std::cerr;
	std::ofstream fout(h.output.c_str());
	if (!fout) die("Unable to open output file for writing");
	// main.cpp:131-147
	fout << "PCRamp version 0.3" << std::endl;
	fout << "Command line:";
	for (int i = 0; i < argc; ++i) fout << ' ' << argv[i];
	fout << std::endl;
	fout << "Random number seed = " << h.seed << std::endl;
	vout << "PCRamp version 0.3 (pcramp_b200: design loop on the B200 path)" << std::endl;
	vout << "Random number seed = " << h.seed << std::endl;

	pcramp_gpu_ctx *ctx = nullptr;
	if (pcramp_gpu_create(&ctx, h.device)) die(ctx ? pcramp_gpu_last_error(ctx) : "no CUDA device (this host has no CPU path)");
	const Loaded targets = load_collection(ctx, PCRAMP_TARGET, h.target_files, std::max(h.d.target_amplicon_min, h.target_len_min), h.target_len_max,
		h.target_ignore, h.normalize_target);
	const Loaded backgrounds = h.background_files.empty() ? Loaded()
		: load_collection(ctx, PCRAMP_BACKGROUND, h.background_files, std::max(h.d.background_amplicon_min, h.background_len_min),
			h.background_len_max, h.background_ignore, h.normalize_background);
	if (targets.n == 0) die("Did not read any target sequences"); // main.cpp:436-438 throws here
	sequence_summary("target sequence summary", targets, fout);
	sequence_summary("Target:", targets, vout);
	sequence_summary("background sequence summary", backgrounds, fout);
	sequence_summary("Background:", backgrounds, vout);

	pcramp_gpu_design *design = nullptr;
	if (pcramp_gpu_design_create(ctx, &h.d, h.seed, &design)) die(pcramp_gpu_last_error(ctx));
	const uint32_t t_words = (targets.n + 31u) / 32u, b_words = (backgrounds.n + 31u) / 32u;
	std::vector<uint32_t> t_bits(std::max(1u, t_words)), b_bits(std::max(1u, b_words));
	const char *rule = "###########################################################################################";
	for (;;) {
		pcramp_gpu_design_result res;
		if (pcramp_gpu_design_iteration(design, &res)) die(pcramp_gpu_design_last_error(design));
		vout << "Design iteration " << res.iteration << std::endl;
		fout << rule << std::endl;
		fout << "# Attempting to detect " << res.targets_remaining << " remaining targets" << std::endl;
		vout << "\t\tNumber of active target sequences = " << res.num_active_target << " (total weight = " << res.active_target_norm << ")" << std::endl;
		vout << "\tTarget word table has " << res.n_target_entries << " entries" << std::endl;
		if (h.timing)
			vout << "\t[timing, ms] candidates " << res.ms_candidates << "  background index " << res.ms_select_background << "  target index "
			     << res.ms_select_target << "  optimize " << res.ms_optimize << "  screen " << res.ms_screen << "  accept " << res.ms_accept << "  total "
			     << res.ms_total << std::endl;
		if (!res.found) break; // main.cpp:928-932
		const std::string fs = word_string(res.f, res.reused_f != 0), rs = word_string(res.r, res.reused_r != 0);
		vout << "\tBest assay: " << word_string(res.f, false) << '\t' << word_string(res.r, false) << "\tD(F)=" << res.degeneracy_f << ";D(R)=" << res.degeneracy_r
		     << std::endl;
		vout << "\tBest accuracy = " << (res.target_coverage - res.background_coverage) << " (" << res.target_coverage << " target, "
		     << res.background_coverage << " background)";
		if (h.d.use_multiplex) vout << "; multiplex overlap = " << res.oligo_overlap;
		vout << std::endl;
		// main.cpp:950-975
		fout << "# Assay " << res.major_id << '.' << res.minor_id << " has target coverage score = " << res.target_coverage << " ("
		     << (res.target_coverage * 100.0f) / res.active_target_norm << "% of active) and background coverage score = " << res.background_coverage
		     << " (" << ((res.num_active_background == 0) ? 0.0f : (res.background_coverage * 100.0f) / res.active_background_norm) << "% of active)"
		     << std::endl;
		fout << "ASSAY." << res.major_id << '.' << res.minor_id << '\t';
		if (h.d.use_multiplex) {
			fout << fs << '\t' << rs << "\tD(F)=" << res.degeneracy_f << ";D(R)=" << res.degeneracy_r << std::endl;
			vout << "\tAdded " << res.n_amplicons_added << " amplicon(s) to the multiplex assay background" << std::endl;
		} else {
			fout << word_string(res.f, false) << '\t' << word_string(res.r, false) << "\tD(F)=" << res.degeneracy_f << ";D(R)=" << res.degeneracy_r
			     << std::endl;
		}
		if (pcramp_gpu_design_matches(design, t_bits.data(), b_bits.data())) die(pcramp_gpu_design_last_error(design));
		for (uint32_t i = 0; i < targets.n; ++i) // :1041-1046
			if ((t_bits[i >> 5] >> (i & 31u)) & 1u) fout << "T-" << targets.defline[i] << std::endl;
		for (uint32_t i = 0; i < backgrounds.n; ++i)
			if ((b_bits[i >> 5] >> (i & 31u)) & 1u) fout << "B-" << backgrounds.defline[i] << std::endl;
		if (res.iteration >= h.num_assay) break; // :1126-1129
	}
	// main.cpp:1131-1192
	std::vector<uint8_t> active(std::max(1u, targets.n));
	if (pcramp_gpu_design_active(design, active.data(), b_bits.data())) die(pcramp_gpu_design_last_error(design));
	unsigned int undetected = 0, cross = 0;
	for (uint32_t i = 0; i < targets.n; ++i) undetected += active[i] ? 1u : 0u;
	for (uint32_t i = 0; i < backgrounds.n; ++i) cross += (b_bits[i >> 5] >> (i & 31u)) & 1u;
	if (undetected == 0) vout << "Detected all targets" << std::endl;
	else vout << "Failed to detect a total of " << undetected << " targets" << std::endl;
	vout << "Cross reacted with a total of " << cross << " background sequences" << std::endl;
	fout << rule << std::endl;
	if (undetected == 0) fout << "# Detected all targets" << std::endl;
	else {
		fout << "# Failed to detect a total of " << undetected << " targets" << std::endl;
		fout << "# The following targets were *not* detected by any assay" << std::endl;
		for (uint32_t i = 0; i < targets.n; ++i)
			if (active[i]) fout << "-T-" << targets.defline[i] << std::endl;
	}
	fout << rule << std::endl;
	fout << "# Cross reacted with a total of " << cross << " background sequences" << std::endl;
	for (uint32_t i = 0; i < backgrounds.n; ++i)
		if ((b_bits[i >> 5] >> (i & 31u)) & 1u) fout << "+B-" << backgrounds.defline[i] << std::endl;
	pcramp_gpu_design_destroy(design);
	pcramp_gpu_destroy(ctx);
	vout << "Finished in " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s" << std::endl;
	return EXIT_SUCCESS;
}
