// inflate_common.cuh — status codes and launchers shared by the two gunzip implementations (inflate.cu, inflate2.cu)
#pragma once
#include "common.cuh"

namespace gams {

// status of one blob (include/gams_cuda.h, gams_cuda_gunzip_batch); INF_ETOKENS is internal: the blob did not fit its token
// budget in the two-phase decoder and is decoded by the warp-per-blob kernel instead
enum { INF_OK = 0, INF_EHEADER = 1, INF_ECODE = 2, INF_EOVERFLOW = 3, INF_EDIST = 4, INF_ESIZE = 5, INF_ETRUNC = 6, INF_ECRC = 7,
       INF_ETOKENS = 100 };

// inflate.cu: one warp per blob; only_status >= 0: only the blobs whose status[] holds that value are decoded
int32_t launch_gunzip_only(cudaStream_t st, const uint8_t *d_gz, const uint64_t *d_gz_off, int64_t n_blob, uint8_t *d_out,
                           const uint64_t *d_out_off, int32_t *d_status, int32_t only_status, uint32_t *d_crc_want);
// inflate2.cu: decoder 0: tokens (one thread per blob) -> bytes (one warp per blob); decoder 1: the kernel of inflate.cu;
// then the CRC-32 pass when asked for.  Scratch comes from pd.scratch[7]
int32_t launch_gunzip2(PerDevice &pd, cudaStream_t st, const uint8_t *d_gz, const uint64_t *d_gz_off, int64_t n_blob, uint8_t *d_out,
                       const uint64_t *d_out_off, int32_t *d_status, uint64_t total_out, int decoder, bool check_crc, int n_arrive = 0,
                       const int64_t *arrive_end = nullptr, const cudaEvent_t *arrive_ev = nullptr, cudaEvent_t decoded_ev = nullptr);
// n_arrive > 0: the blobs are still being uploaded; arrive_ev[i] fires once blobs < arrive_end[i] (rising) are on the device, and
// the caller has NOT made `st` wait for the upload.  The header search follows the upload stretch by stretch.
// decoded_ev: recorded on `st` once the output bytes are final, before the CRC-32 pass (which only reads them and sets status codes).
// gunzip2_scratch_bytes: what a batch of (n_blob, total_out) takes of pd.scratch[7]; every piece of the batch takes less, so a
// caller that decodes a batch piece by piece on one stream reserves this first and no piece re-allocates the buffer.
int32_t gunzip2_scratch_bytes(PerDevice &pd, int64_t n_blob, uint64_t total_out, int decoder, bool check_crc, size_t *bytes);

}  // namespace gams
