// experiment: tcgen05.mma kind::f16 with A = fp16 and B = bf16 (mixed operand formats)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <vector>
#include <cmath>
__device__ __forceinline__ uint32_t smem_u32(const void* p){return (uint32_t)__cvta_generic_to_shared(p);}
__device__ __forceinline__ uint64_t desc(uint32_t a){return (uint64_t)((a&0x3FFFFu)>>4)|((uint64_t)1<<16)|((uint64_t)(1024>>4)<<32)|((uint64_t)1<<46)|((uint64_t)2<<61);}
__global__ void k(const __half* A, const __nv_bfloat16* B, float* D, int afmt, int bfmt){
  __shared__ __align__(1024) uint8_t a_s[128*128];
  __shared__ __align__(1024) uint8_t b_s[64*128];
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  int tid=threadIdx.x;
  // A: [128][64] K-major, 128B rows, swizzle
  for(int i=tid;i<128*8;i+=128){int r=i>>3,j=i&7; *(uint4*)(a_s+r*128+((j^(r&7))<<4))=*(const uint4*)((const uint8_t*)A+r*128+j*16);}
  for(int i=tid;i<64*8;i+=128){int r=i>>3,j=i&7; *(uint4*)(b_s+r*128+((j^(r&7))<<4))=*(const uint4*)((const uint8_t*)B+r*128+j*16);}
  if(tid==0){asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(smem_u32(&bar)));asm volatile("fence.mbarrier_init.release.cluster;":::"memory");}
  if(tid<32){asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;"::"r"(smem_u32(&slot)):"memory");asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;":::"memory");}
  asm volatile("fence.proxy.async.shared::cta;":::"memory");
  asm volatile("tcgen05.fence::before_thread_sync;":::"memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;":::"memory");
  uint32_t tm=slot;
  uint32_t idesc=(1u<<4)|((uint32_t)afmt<<7)|((uint32_t)bfmt<<10)|((64u>>3)<<17)|((128u>>4)<<24);
  if(tid==0){
    uint64_t da=desc(smem_u32(a_s)), db=desc(smem_u32(b_s));
    for(int kk=0;kk<4;++kk){uint32_t acc=kk!=0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"::"r"(tm),"l"(da+(uint64_t)(kk*2)),"l"(db+(uint64_t)(kk*2)),"r"(idesc),"r"(acc):"memory");}
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"::"r"(smem_u32(&bar)):"memory");
  }
  uint32_t done=0; while(!done){asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}":"=r"(done):"r"(smem_u32(&bar)),"r"(0):"memory");}
  asm volatile("tcgen05.fence::after_thread_sync;":::"memory");
  int warp=tid>>5;
  for(int cc=0;cc<2;++cc){uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];":"=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),"=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]):"r"(tm+cc*32+((uint32_t)(warp*32)<<16)):"memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;":::"memory");
    for(int i=0;i<32;++i) D[tid*64+cc*32+i]=__uint_as_float(r[i]);}
  asm volatile("tcgen05.fence::before_thread_sync;":::"memory"); __syncthreads();
  if(tid<32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;"::"r"(tm):"memory");
}
int main(){
  std::vector<__half> A(128*64); std::vector<__nv_bfloat16> B(64*64); std::vector<float> Af(128*64),Bf(64*64);
  srand(1); for(int i=0;i<128*64;++i){float v=(rand()%2001-1000)/37.f; A[i]=__float2half(v); Af[i]=__half2float(A[i]);}
  for(int i=0;i<64*64;++i){float v=(rand()%2001-1000)/911.f; B[i]=__float2bfloat16(v); Bf[i]=__bfloat162float(B[i]);}
  __half* dA; __nv_bfloat16* dB; float* dD; cudaMalloc(&dA,128*64*2); cudaMalloc(&dB,64*64*2); cudaMalloc(&dD,128*64*4);
  cudaMemcpy(dA,A.data(),128*64*2,cudaMemcpyHostToDevice); cudaMemcpy(dB,B.data(),64*64*2,cudaMemcpyHostToDevice);
  k<<<1,128>>>(dA,dB,dD,0,1); cudaError_t e=cudaDeviceSynchronize(); printf("mixed f16 x bf16: %s\n",cudaGetErrorString(e));
  std::vector<float> D(128*64); cudaMemcpy(D.data(),dD,128*64*4,cudaMemcpyDeviceToHost);
  double maxerr=0,maxref=0; for(int m=0;m<128;++m)for(int n=0;n<64;++n){double s=0;for(int kk=0;kk<64;++kk)s+=(double)Af[m*64+kk]*Bf[n*64+kk]; maxerr=fmax(maxerr,fabs(s-D[m*64+n])); maxref=fmax(maxref,fabs(s));}
  printf("max abs err %.4g (max ref %.4g)\n",maxerr,maxref); return 0;}
