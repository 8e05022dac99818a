#include <cstdint>
struct u64x4 { uint64_t a,b,c,d; };
__device__ __forceinline__ u64x4 ld256(const void* p){ u64x4 v; asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(v.a),"=l"(v.b),"=l"(v.c),"=l"(v.d):"l"(p)); return v;}
__device__ __forceinline__ void st256(void* p, u64x4 v){ asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};"::"l"(p),"l"(v.a),"l"(v.b),"l"(v.c),"l"(v.d):"memory");}
__device__ __forceinline__ uint64_t pol(){ uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;":"=l"(p)); return p;}
__device__ __forceinline__ ulonglong2 ld128(const void* p, uint64_t po){ ulonglong2 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;":"=l"(v.x),"=l"(v.y):"l"(p),"l"(po)); return v;}
__global__ void k(const uint64_t* a, uint64_t* o, size_t n){ size_t i=blockIdx.x*blockDim.x+threadIdx.x; uint64_t po=pol(); u64x4 v=ld256(a+4*i); ulonglong2 w=ld128(a+2*i,po); v.a+=w.x; v.b+=w.y; st256(o+4*i,v);}
