/* Single-rank MPI stand-in used ONLY to build the unmodified reference as the
 * parity oracle (oracle/_ref).  TEST INFRASTRUCTURE - never linked into the
 * product library.
 *
 * The reference exchanges halos with itself when run on one rank
 * (reference src/atm/Connectivity.cpp:941 MPI_Irecv, :971 MPI_Isend,
 * :1081 MPI_Test), so Isend must deliver into the oldest posted, unmatched
 * Irecv on the same rank and MPI_Test must then report completion.
 */
#ifndef TB200_ORACLE_MPI_STUB_H
#define TB200_ORACLE_MPI_STUB_H

#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <deque>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;

struct MPI_Status {
	int MPI_SOURCE;
	int MPI_TAG;
	int MPI_ERROR;
};

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ERR_IN_STATUS 17

#define MPI_BYTE   1
#define MPI_CHAR   2
#define MPI_INT    3
#define MPI_LONG   4
#define MPI_DOUBLE 5

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IN_PLACE ((void*)(-1))
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

namespace tb200_mpi_stub {

struct PostedRecv {
	void * buf;
	size_t bytes;
	bool done;
};

inline std::deque<PostedRecv> & Posted() {
	static std::deque<PostedRecv> q;
	return q;
}

inline size_t TypeSize(MPI_Datatype t) {
	switch (t) {
		case MPI_BYTE: return 1;
		case MPI_CHAR: return 1;
		case MPI_INT: return sizeof(int);
		case MPI_LONG: return sizeof(long);
		case MPI_DOUBLE: return sizeof(double);
	}
	std::fprintf(stderr, "mpi stub: unknown datatype %d\n", t);
	std::abort();
}

}  // namespace

inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Abort(MPI_Comm, int code) { std::exit(code); return 0; }
inline int MPI_Comm_rank(MPI_Comm, int * r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int * s) { *s = 1; return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }

inline int MPI_Reduce(
	const void * send, void * recv, int count, MPI_Datatype t,
	MPI_Op, int, MPI_Comm
) {
	if (send != MPI_IN_PLACE) {
		std::memcpy(recv, send, count * tb200_mpi_stub::TypeSize(t));
	}
	return 0;
}

inline int MPI_Allreduce(
	const void * send, void * recv, int count, MPI_Datatype t,
	MPI_Op, MPI_Comm
) {
	if (send != MPI_IN_PLACE) {
		std::memcpy(recv, send, count * tb200_mpi_stub::TypeSize(t));
	}
	return 0;
}

inline int MPI_Irecv(
	void * buf, int count, MPI_Datatype t, int, int, MPI_Comm,
	MPI_Request * req
) {
	using namespace tb200_mpi_stub;
	PostedRecv r;
	r.buf = buf;
	r.bytes = count * TypeSize(t);
	r.done = false;
	Posted().push_back(r);
	*req = (int)(Posted().size()) - 1;
	return 0;
}

inline int MPI_Isend(
	const void * buf, int count, MPI_Datatype t, int, int, MPI_Comm,
	MPI_Request * req
) {
	using namespace tb200_mpi_stub;
	size_t bytes = count * TypeSize(t);
	for (size_t i = 0; i < Posted().size(); i++) {
		if (!Posted()[i].done) {
			if (Posted()[i].bytes < bytes) {
				std::fprintf(stderr, "mpi stub: message truncated\n");
				std::abort();
			}
			std::memcpy(Posted()[i].buf, buf, bytes);
			Posted()[i].done = true;
			*req = -1;
			return 0;
		}
	}
	std::fprintf(stderr, "mpi stub: Isend without a posted Irecv\n");
	std::abort();
}

inline int MPI_Recv(
	void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *
) {
	std::fprintf(stderr, "mpi stub: blocking MPI_Recv on one rank\n");
	std::abort();
}

inline int MPI_Test(MPI_Request * req, int * flag, MPI_Status * st) {
	using namespace tb200_mpi_stub;
	if (*req < 0 || *req >= (int)Posted().size()) {
		*flag = 1;
	} else {
		*flag = Posted()[*req].done ? 1 : 0;
	}
	if (st != 0) {
		st->MPI_SOURCE = 0;
		st->MPI_TAG = 0;
		st->MPI_ERROR = 0;
	}
	if (*flag) {
		// retire the whole queue once every posted receive is complete
		bool all = true;
		for (size_t i = 0; i < Posted().size(); i++) {
			all = all && Posted()[i].done;
		}
		if (all) Posted().clear();
	}
	return 0;
}

inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }

#endif
