// nka_vec.cu -- device vectors with the operations of the reference's abstract `vector`
// class (src-F08-vector/vector_class.F90:90-109), i.e. what a concrete `gpu_vector`
// extension needs: clone, copy, setval, scale, the four update forms, dot, norm2.
// Expression order follows the reference's concrete grid_vector
// (src-F08-vector/grid_vector_type.F90:108-197).  Reductions are deterministic
// (fixed-order two-stage) and, with a communicator, summed over the ranks' slabs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/nka_b200.h"
#include "nka_internal.h"
#include "nka_kernels.cuh"

struct nka_vec {
  size_t n = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  double* d = nullptr;
  double* partials = nullptr;   // device, grid x 1
  unsigned* ticket = nullptr;   // device
  double* result = nullptr;     // device, 1
  double* result_host = nullptr;  // pinned, 1
  int num_sms = 0;
  NkaComm* comm = nullptr;
};

void nka_attach_shared_comm(NKA st, NkaComm* c);   // nka_capi.cu

static int vec_grid(const nka_vec* v, size_t work)
{
  size_t need = (work + NKA_THREADS - 1) / NKA_THREADS;
  size_t cap = (size_t)v->num_sms * 8;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- element-wise kernels: double2 bulk + scalar tail -------------------------------------
// form 0: z = val          1: z = a*z           2: z = a*x + z        3: z = a*x + b*z
// form 4: z = a*x + b*y + z        5: z = a*x + b*y + c*z
template <int FORM>
__device__ __forceinline__ double vec_apply(double z, double x, double y, double a, double b, double c)
{
  if (FORM == 0) return a;
  if (FORM == 1) return a * z;
  if (FORM == 2) return a * x + z;
  if (FORM == 3) return a * x + b * z;
  if (FORM == 4) return a * x + b * y + z;
  return a * x + b * y + c * z;
}

template <int FORM>
__global__ void __launch_bounds__(NKA_THREADS)
nka_vec_elementwise(double* __restrict__ z, const double* __restrict__ x, const double* __restrict__ y,
                    double a, double b, double c, size_t n)
{
  const size_t stride = (size_t)gridDim.x * NKA_THREADS;
  const size_t start = (size_t)blockIdx.x * NKA_THREADS + threadIdx.x;
  const size_t nv = n / 2;
  double2* z2 = reinterpret_cast<double2*>(z);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const double2* y2 = reinterpret_cast<const double2*>(y);
  for (size_t i = start; i < nv; i += stride) {
    double2 zv = (FORM >= 1) ? z2[i] : make_double2(0.0, 0.0);
    const double2 xv = (FORM >= 2) ? __ldg(x2 + i) : make_double2(0.0, 0.0);
    const double2 yv = (FORM >= 4) ? __ldg(y2 + i) : make_double2(0.0, 0.0);
    zv.x = vec_apply<FORM>(zv.x, xv.x, yv.x, a, b, c);
    zv.y = vec_apply<FORM>(zv.y, xv.y, yv.y, a, b, c);
    z2[i] = zv;
  }
  if ((n & 1) && start == 0) {
    const size_t i = n - 1;
    z[i] = vec_apply<FORM>(FORM >= 1 ? z[i] : 0.0, FORM >= 2 ? x[i] : 0.0, FORM >= 4 ? y[i] : 0.0, a, b, c);
  }
}

__global__ void __launch_bounds__(NKA_THREADS)
nka_vec_dot_kernel(const double* __restrict__ x, const double* __restrict__ y, size_t n,
                   double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ out)
{
  const size_t stride = (size_t)gridDim.x * NKA_THREADS;
  const size_t start = (size_t)blockIdx.x * NKA_THREADS + threadIdx.x;
  const size_t nv = n / 2;
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const double2* y2 = reinterpret_cast<const double2*>(y);
  double acc[1] = {0.0};
  for (size_t i = start; i < nv; i += stride) {
    const double2 a = __ldg(x2 + i), b = __ldg(y2 + i);
    acc[0] = fma(a.x, b.x, acc[0]);
    acc[0] = fma(a.y, b.y, acc[0]);
  }
  if ((n & 1) && start == 0) acc[0] = fma(x[n - 1], y[n - 1], acc[0]);
  nka_grid_reduce<1, NKA_THREADS>(acc, partials, ticket, partials, gridDim.x, [&](int, double v) { out[0] = v; });
}

// ---- life cycle ------------------------------------------------------------------------
extern "C" NKAVEC nka_vec_create(size_t n, int device, void* stream)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    nka_fail(__FILE__, __LINE__, "no CUDA device: libnka_b200 has no CPU compute path");
  nka_vec* v = new nka_vec();
  if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
  v->device = device;
  DeviceGuard guard(device);
  CUDA_CHECK(cudaDeviceGetAttribute(&v->num_sms, cudaDevAttrMultiProcessorCount, device));
  v->n = n;
  v->stream = (cudaStream_t)stream;
  const size_t ld = ((n + 15) / 16) * 16;
  CUDA_CHECK(cudaMalloc(&v->d, (ld ? ld : 16) * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&v->partials, (size_t)v->num_sms * 8 * sizeof(double)));
  CUDA_CHECK(cudaMalloc(&v->ticket, sizeof(unsigned)));
  CUDA_CHECK(cudaMalloc(&v->result, sizeof(double)));
  CUDA_CHECK(cudaMallocHost(&v->result_host, sizeof(double)));
  CUDA_CHECK(cudaMemsetAsync(v->ticket, 0, sizeof(unsigned), v->stream));
  return v;
}

extern "C" NKAVEC nka_vec_clone(NKAVEC src)
{
  // allocate(clone, source=this): src-F08-vector/grid_vector_type.F90:86-97
  NKA_REQUIRE(src != NULL, "nka_vec_clone: null vector");
  NKAVEC v = nka_vec_create(src->n, src->device, (void*)src->stream);
  v->comm = nka_comm_retain(src->comm);
  DeviceGuard guard(src->device);
  CUDA_CHECK(cudaMemcpyAsync(v->d, src->d, src->n * sizeof(double), cudaMemcpyDeviceToDevice, src->stream));
  return v;
}

extern "C" void nka_vec_destroy(NKAVEC v)
{
  if (!v) return;
  DeviceGuard guard(v->device);
  cudaStreamSynchronize(v->stream);
  cudaFree(v->d); cudaFree(v->partials); cudaFree(v->ticket); cudaFree(v->result);
  cudaFreeHost(v->result_host);
  nka_comm_release(v->comm);
  delete v;
}

extern "C" size_t nka_vec_size(NKAVEC v) { NKA_REQUIRE(v != NULL, "nka_vec_size: null vector"); return v->n; }
extern "C" double* nka_vec_data(NKAVEC v) { NKA_REQUIRE(v != NULL, "nka_vec_data: null vector"); return v->d; }

extern "C" void nka_vec_set_host(NKAVEC v, const double* host)
{
  NKA_REQUIRE(v != NULL, "nka_vec_set_host: null vector");
  DeviceGuard guard(v->device);
  CUDA_CHECK(cudaMemcpyAsync(v->d, host, v->n * sizeof(double), cudaMemcpyHostToDevice, v->stream));
  CUDA_CHECK(cudaStreamSynchronize(v->stream));
}

extern "C" void nka_vec_get_host(NKAVEC v, double* host)
{
  NKA_REQUIRE(v != NULL, "nka_vec_get_host: null vector");
  DeviceGuard guard(v->device);
  CUDA_CHECK(cudaMemcpyAsync(host, v->d, v->n * sizeof(double), cudaMemcpyDeviceToHost, v->stream));
  CUDA_CHECK(cudaStreamSynchronize(v->stream));
}

static void check_same(NKAVEC a, NKAVEC b, const char* what)
{
  // vector_class.F90:151-181: mismatched operands are an error stop
  if (!a || !b || a->n != b->n || a->device != b->device) nka_fail(__FILE__, __LINE__, what);
}

template <int FORM>
static void launch_elementwise(NKAVEC z, NKAVEC x, NKAVEC y, double a, double b, double c)
{
  DeviceGuard guard(z->device);
  nka_vec_elementwise<FORM><<<vec_grid(z, z->n / 2 + 1), NKA_THREADS, 0, z->stream>>>(
      z->d, x ? x->d : nullptr, y ? y->d : nullptr, a, b, c, z->n);
  CUDA_CHECK(cudaGetLastError());
}

extern "C" void nka_vec_copy(NKAVEC dst, NKAVEC src)
{
  check_same(dst, src, "incompatible arguments to nka_vec_copy");
  DeviceGuard guard(dst->device);
  CUDA_CHECK(cudaMemcpyAsync(dst->d, src->d, dst->n * sizeof(double), cudaMemcpyDeviceToDevice, dst->stream));
}

extern "C" void nka_vec_setval(NKAVEC v, double val)
{
  NKA_REQUIRE(v != NULL, "nka_vec_setval: null vector");
  launch_elementwise<0>(v, nullptr, nullptr, val, 0.0, 0.0);
}

extern "C" void nka_vec_scale(NKAVEC v, double a)
{
  NKA_REQUIRE(v != NULL, "nka_vec_scale: null vector");
  launch_elementwise<1>(v, nullptr, nullptr, a, 0.0, 0.0);
}

// The zero-coefficient short cuts of the base class (vector_class.F90:176,189,203-206,219-222)
// belong to the caller's non-virtual wrappers; these are the deferred *_ procedures.
extern "C" void nka_vec_update1(NKAVEC y, double a, NKAVEC x)
{
  check_same(y, x, "incompatible arguments to nka_vec_update1");
  launch_elementwise<2>(y, x, nullptr, a, 0.0, 0.0);
}

extern "C" void nka_vec_update2(NKAVEC y, double a, NKAVEC x, double b)
{
  check_same(y, x, "incompatible arguments to nka_vec_update2");
  launch_elementwise<3>(y, x, nullptr, a, b, 0.0);
}

extern "C" void nka_vec_update3(NKAVEC z, double a, NKAVEC x, double b, NKAVEC y)
{
  check_same(z, x, "incompatible arguments to nka_vec_update3");
  check_same(z, y, "incompatible arguments to nka_vec_update3");
  launch_elementwise<4>(z, x, y, a, b, 0.0);
}

extern "C" void nka_vec_update4(NKAVEC z, double a, NKAVEC x, double b, NKAVEC y, double c)
{
  check_same(z, x, "incompatible arguments to nka_vec_update4");
  check_same(z, y, "incompatible arguments to nka_vec_update4");
  launch_elementwise<5>(z, x, y, a, b, c);
}

extern "C" double nka_vec_dot(NKAVEC x, NKAVEC y)
{
  check_same(x, y, "incompatible arguments to nka_vec_dot");
  DeviceGuard guard(x->device);
  nka_vec_dot_kernel<<<vec_grid(x, x->n / 2 + 1), NKA_THREADS, 0, x->stream>>>(x->d, y->d, x->n, x->partials,
                                                                              x->ticket, x->result);
  CUDA_CHECK(cudaGetLastError());
  if (x->comm) {
    const int rc = g_nccl.AllReduce(x->result, x->result, 1, kNcclFloat64, kNcclSum, x->comm->comm, x->stream);
    if (rc != 0) nka_fail(__FILE__, __LINE__, "ncclAllReduce failed in nka_vec_dot");
  }
  CUDA_CHECK(cudaMemcpyAsync(x->result_host, x->result, sizeof(double), cudaMemcpyDeviceToHost, x->stream));
  CUDA_CHECK(cudaStreamSynchronize(x->stream));
  return *x->result_host;
}

extern "C" double nka_vec_norm2(NKAVEC x)
{
  return sqrt(nka_vec_dot(x, x));
}

extern "C" int nka_vec_comm_init(NKAVEC v, int nranks, int rank, const void* id128)
{
  NKA_REQUIRE(v != NULL && id128 != NULL, "nka_vec_comm_init: null argument");
  NKA_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "nka_vec_comm_init: bad rank/nranks");
  if (!nka_nccl_load()) return -1;
  DeviceGuard guard(v->device);
  NkaId128 id;
  memcpy(id.bytes, id128, sizeof id.bytes);
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (rc != 0) return rc;
  nka_comm_release(v->comm);
  v->comm = new NkaComm();
  v->comm->comm = comm; v->comm->owned = true; v->comm->nranks = nranks; v->comm->rank = rank;
  return 0;
}

// init(vec, mvec) of the vector flavour (src-F08-vector/nka_type.F90:175-188): an accelerator
// shaped like `proto` (length, device, stream, communicator).
extern "C" NKA nka_init_like(NKAVEC proto, int mvec, double vtol)
{
  NKA_REQUIRE(proto != NULL, "nka_init_like: null vector");
  NKA st = nka_init_ex(proto->n, mvec, vtol, proto->device, (void*)proto->stream);
  if (proto->comm) nka_attach_shared_comm(st, proto->comm);
  return st;
}

// accel_update(f) with class(vector) f (src-F08-vector/nka_type.F90:219-222)
extern "C" void nka_accel_update_vec(NKA st, NKAVEC f)
{
  NKA_REQUIRE(st != NULL && f != NULL, "nka_accel_update_vec: null argument");
  NKA_REQUIRE(nka_vec_len64(st) == f->n, "nka_accel_update_vec: vector length differs from the accelerator's");
  nka_accel_update_dev(st, f->d);
}
