/* mpi.h — minimal single-node MPI for building the UNMODIFIED reference (helmholtz-analytics/mpi4torch)
 * on a box without an MPI installation.  Only the subset of MPI-3 the reference calls
 *  plus a few conveniences.  Transport: POSIX shared memory + C++11
 * atomics (libmpi.so from ../src/mpi_shim.cpp).  Handles are plain ints (MPICH-style), so the
 * Fortran conversions are identities.  This is NOT CUDA-aware: like a stock distro OpenMPI/MPICH
 * it only accepts host pointers, so the reference takes its own host-staging path
 * (reference csrc/extension.cpp:61-104). */
#ifndef MPISHIM_MPI_H
#define MPISHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPI_VERSION 3
#define MPI_SUBVERSION 1
#define MPISHIM 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Fint;
typedef long MPI_Aint;
typedef long long MPI_Count;
typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  long long shim_bytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_ERR_ARG 12
#define MPI_ERR_TYPE 3
#define MPI_ERR_OP 9
#define MPI_ERR_TRUNCATE 14
#define MPI_ERR_REQUEST 19

#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_COMM_WORLD ((MPI_Comm)1)
#define MPI_COMM_SELF ((MPI_Comm)2)

#define MPI_DATATYPE_NULL ((MPI_Datatype)0)
#define MPI_BYTE ((MPI_Datatype)1)
#define MPI_CHAR ((MPI_Datatype)2)
#define MPI_SIGNED_CHAR ((MPI_Datatype)2)
#define MPI_SHORT ((MPI_Datatype)3)
#define MPI_INT ((MPI_Datatype)4)
#define MPI_LONG ((MPI_Datatype)5)
#define MPI_LONG_LONG ((MPI_Datatype)5)
#define MPI_FLOAT ((MPI_Datatype)6)
#define MPI_DOUBLE ((MPI_Datatype)7)
#define MPI_UNSIGNED_CHAR ((MPI_Datatype)8)
#define MPI_UNSIGNED ((MPI_Datatype)9)
#define MPI_UNSIGNED_LONG ((MPI_Datatype)10)
#define MPI_INT8_T ((MPI_Datatype)2)
#define MPI_INT16_T ((MPI_Datatype)3)
#define MPI_INT32_T ((MPI_Datatype)4)
#define MPI_INT64_T ((MPI_Datatype)5)
#define MPI_UINT8_T ((MPI_Datatype)8)

#define MPI_OP_NULL ((MPI_Op)0)
#define MPI_MAX ((MPI_Op)1)
#define MPI_MIN ((MPI_Op)2)
#define MPI_SUM ((MPI_Op)3)
#define MPI_PROD ((MPI_Op)4)
#define MPI_LAND ((MPI_Op)5)
#define MPI_BAND ((MPI_Op)6)
#define MPI_LOR ((MPI_Op)7)
#define MPI_BOR ((MPI_Op)8)
#define MPI_LXOR ((MPI_Op)9)
#define MPI_BXOR ((MPI_Op)10)
#define MPI_MINLOC ((MPI_Op)11)
#define MPI_MAXLOC ((MPI_Op)12)

#define MPI_REQUEST_NULL ((MPI_Request)0)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_BOTTOM ((void*)0)
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-2)
#define MPI_PROC_NULL (-3)
#define MPI_UNDEFINED (-32766)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_PROCESSOR_NAME 256

#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

int MPI_Init(int* argc, char*** argv);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Initialized(int* flag);
int MPI_Finalized(int* flag);
int MPI_Finalize(void);
int MPI_Query_thread(int* provided);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
int MPI_Error_string(int errorcode, char* string, int* resultlen);
int MPI_Get_processor_name(char* name, int* resultlen);

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
MPI_Comm MPI_Comm_f2c(MPI_Fint comm);
MPI_Fint MPI_Comm_c2f(MPI_Comm comm);
MPI_Request MPI_Request_f2c(MPI_Fint request);
MPI_Fint MPI_Request_c2f(MPI_Request request);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
               int root, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                const int* displs, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Allgatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, const int* recvcounts,
                   const int* displs, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Scatter(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                int root, MPI_Comm comm);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs, MPI_Datatype sendtype, void* recvbuf,
                 int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);

int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Wait(MPI_Request* request, MPI_Status* status);
int MPI_Test(MPI_Request* request, int* flag, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request requests[], MPI_Status statuses[]);
int MPI_Get_count(const MPI_Status* status, MPI_Datatype datatype, int* count);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_create_resized(MPI_Datatype oldtype, MPI_Aint lb, MPI_Aint extent, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* datatype);
int MPI_Type_free(MPI_Datatype* datatype);
int MPI_Type_get_extent(MPI_Datatype datatype, MPI_Aint* lb, MPI_Aint* extent);
int MPI_Type_size(MPI_Datatype datatype, int* size);

#ifdef __cplusplus
}
#endif
#endif /* MPISHIM_MPI_H */
