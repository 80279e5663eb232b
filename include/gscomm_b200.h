/*
 * gscomm_b200.h — the one exchange step of the keyframe-sharded mapping step: the SUM all-reduce of
 * the flat fp32 gradient bucket over the GPUs of one NVLink / NVSwitch node, as ONE kernel over peer memory.
 *
 * The reference has no multi-GPU path (its mapper optimises one keyframe per step, R/slam/mapper.py:797-939); this is
 * the collective that the K-keyframe construct adds.  Every rank passes the device addresses at which it sees the
 * bucket of every rank (peer mappings of one symmetric allocation: CUDA IPC / cuMem handles; the host side obtains
 * them from torch.distributed._symmetric_memory) and, when the node has NVLS, the multicast address of the same
 * allocation.  Rank r owns the r-th slice of the bucket:
 *
 *   peer mode : r loads its slice from all ranks (P2P loads over NVLink), adds the copies in rank order and stores the
 *               sum into every rank's bucket (P2P stores);
 *   NVLS mode : one multimem.ld_reduce per 16 bytes pulls the sum from the switch, one multimem.st broadcasts it.
 *
 * Either way every element is reduced exactly once, by its owner, and the same bits land on every rank — replicas
 * of the parameters stay identical.  The kernel starts with a flag handshake (every rank's bucket is complete) and
 * ends with one (every rank's stores have landed): no host synchronisation, no NCCL call on the path.
 *
 * flags: nranks device addresses of a zero-initialised uint32 array of gsr_allreduce_flag_words() entries inside the
 * symmetric allocation (one array per rank, same offset everywhere).  epoch: a counter the caller increments by one for
 * every call, the same on all ranks, starting at 1.
 */
#ifndef GSCOMM_B200_H_
#define GSCOMM_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "gsrast_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_COMM_MAX_RANKS 8

size_t gsr_allreduce_flag_words(void);

/* bucket[r] / flags[r]: address of rank r's bucket / flag array as mapped into THIS process (HOST arrays of nranks
 * device pointers; entry [rank] is the local one).  multicast: NVLS address of the bucket or NULL.  n: floats in the
 * bucket (the allocation must be 16-byte aligned; n need not be a multiple of 4). */
int gsr_allreduce_sum_f32(gsr_stream_t stream, int32_t nranks, int32_t rank, float* const* bucket, float* multicast,
                          uint32_t* const* flags, int64_t n, uint32_t epoch);

/* All-gather over the same kind of symmetric buffer: rank r's slice — floats [r * ceil(n4 / nranks) * 4, ...) with
 * n4 = ceil(n / 4), the slicing of the all-reduce — is copied from its own buffer into every other rank's (P2P stores, or
 * one multimem.st per 16 bytes).  Used to replicate the Gaussian parameters when every rank uploads only its 1/nranks
 * share from host memory.  Same flag / epoch rules (a separate flag array and epoch counter per buffer). */
int gsr_allgather_f32(gsr_stream_t stream, int32_t nranks, int32_t rank, float* const* buffer, float* multicast,
                      uint32_t* const* flags, int64_t n, uint32_t epoch);

#ifdef __cplusplus
}
#endif
#endif /* GSCOMM_B200_H_ */
