// fastx.h - host-side FASTA/FASTQ reader with the record semantics of the reference's parser
// (kseq.h:192-232), restated around a large block buffer; plain or gzip input through zlib.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <zlib.h>

namespace yakb {

class FastxReader {
public:
	FastxReader() {}
	~FastxReader() { close(); }
	bool open(const char *fn); // NULL or "-" = stdin (count.c:151)
	void close();
	// next record: sequence bytes (line ends removed) in seq(); returns length, -1 at EOF,
	// -2 on a truncated quality string (kseq.h:189-191)
	int64_t next();
	const std::string &seq() const { return seq_; }
	const std::string &name() const { return name_; }

private:
	int getc_();
	// append the rest of the current line to s (without the '\n'); false if nothing was left
	bool line_(std::string &s, int64_t *count_only);
	gzFile fp_ = nullptr;
	std::vector<unsigned char> buf_;
	int64_t beg_ = 0, end_ = 0;
	bool eof_ = false;
	int last_ = 0, last_qual_ = 0;
	std::string seq_, name_;
};

} // namespace yakb
