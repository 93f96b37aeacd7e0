// Unit-to-unit transport: replaces MDManager's hand-off between units of one process
// (mdmanager.cc:205-217, 305-318) and its MPI_Sendrecv / MPI_Allreduce calls between processes
// (mdmanager.cc:221-229, 321-329, 364, 375, 386).
#pragma once
#include "unit.h"

namespace mdb {

int comm_unique_id(void *out128);
int comm_init(Unit **us, int n, int nranks, int rank, const void *unique_id);
int comm_release(Unit *u);
// migration hop in direction dir: leavers packed in us[i]->mig_send_* (n_leave[i] of them)
int comm_migrate(Unit **us, int n, int dir, const long long *n_leave);
// rebuild-time ghost exchange in direction dir: sizes the receive windows, first transfer
int comm_border_setup(Unit **us, int n, int dir);
// per-step ghost refresh in direction dir (fixed windows)
int comm_border_refresh(Unit **us, int n, int dir, bool on_comm_stream = false);
// Peer-to-peer ghost refresh between processes (comm.cu): comm_p2p_update exchanges the ghost windows after the
// border set-up of a rebuild, comm_refresh_begin opens one refresh round (acknowledges the previous one to the
// senders), comm_p2p_invalidate sends the next refreshes through NCCL until the windows have been exchanged again.
int comm_p2p_update(Unit **us, int n);
int comm_refresh_begin(Unit **us, int n, bool on_comm_stream);
void comm_p2p_invalidate();
bool comm_p2p_enabled();
bool comm_multi_process();
int comm_remote_directions(const Unit *u);  // how many of the 6 neighbours live in another process
int comm_allreduce_or(Unit **us, int n, int *flag);
// MAX over the processes of {flag, bits}: the expiry flag and the bit pattern of max |p|^2 in one collective
int comm_allreduce_flag_bits(Unit **us, int n, int *flag, int *err, unsigned long long *bits);
int comm_allreduce_sum(Unit **us, int n, double *v, int cnt);

int ghost_put_to(Unit *s, int dir, double4 *dst, double4 *dst2, cudaStream_t st);
int observe_kinetic(Unit *u, double out[4]);
int expiry_check_launch(Unit *u);

}  // namespace mdb
