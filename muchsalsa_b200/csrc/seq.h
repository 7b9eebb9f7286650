// seq.h — resident sequence store + batched slice / reverse-complement (kernel K8), internal interface.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/msgpu.h"

namespace msgpu {

struct SeqStore;

struct MappedFile {
  const char *data = nullptr;
  size_t      size = 0;
  bool        ok   = false;
  explicit MappedFile(const char *path);
  ~MappedFile();
  MappedFile(MappedFile const &)            = delete;
  MappedFile &operator=(MappedFile const &) = delete;
};

// SequenceAccessor.cpp:71-78: anything but .fa / .fasta (case-insensitive) is FASTQ
bool path_is_fastq(const char *path);

SeqStore *seq_create(cudaStream_t st, const char *unitigs, size_t n_u, const char *nanopore, size_t n_n,
                     bool nanopore_is_fastq, const std::vector<std::string> &prime_reads,
                     const std::vector<std::string> &prime_unitigs);
void      seq_destroy(SeqStore *s);
bool      seq_lookup(const SeqStore *s, uint32_t which, const char *name, uint32_t *id);
uint64_t  seq_batch_bound(const SeqStore *s, const msgpu_seq_req *reqs, size_t n);
void      seq_fetch_host(SeqStore *s, const msgpu_seq_req *reqs, size_t n, char *out, uint64_t cap, uint64_t *out_off,
                         uint8_t *status);
void      seq_fetch_device(SeqStore *s, cudaStream_t st, const msgpu_seq_req *d_reqs, size_t n, char *d_out,
                           uint64_t cap, uint64_t *d_out_off, uint8_t *d_status);
void      slice_revcomp_device(cudaStream_t st, const char *d_text, const uint64_t *d_text_off, const int32_t *d_left,
                               const int32_t *d_right, const uint8_t *d_revcomp, size_t n, char *d_out, uint64_t cap,
                               uint64_t *d_out_off, uint8_t *d_status);

} // namespace msgpu
