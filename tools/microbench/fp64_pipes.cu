// Micro-benchmark (B200, sm_100a): FP64 pipe throughput for DFMA vs DMMA shapes, and
// read-only HBM streaming bandwidth (LDG.128 vs cp.async.bulk).  Decides K1's inner product path.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
  #pragma unroll
  for (int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0; 
  #pragma unroll
  for (int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void dmma1684(double* c,const double* a,double b){
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n":"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(b));
}
__device__ __forceinline__ void dmma1688(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n":"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n":"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}

template<int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double* out,int iters,double a,double b){
  double c[ILP][2];
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma884(c[i][0],c[i][1],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP,int KK>
__global__ void __launch_bounds__(256) k_dmma16(double* out,int iters,double a0,double b0){
  double c[ILP][4]; double a[8]; double b[4];
  #pragma unroll
  for(int i=0;i<8;i++) a[i]=a0+i*1e-9;
  #pragma unroll
  for(int i=0;i<4;i++) b[i]=b0+i*1e-9;
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x;c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<ILP;i++){
      if(KK==4) dmma1684(c[i],a,b[0]);
      else if(KK==8) dmma1688(c[i],a,b);
      else dmma16816(c[i],a,b);
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// ---- streaming read bandwidth -------------------------------------------------------------
__global__ void __launch_bounds__(256) k_read_ldg(const double2* __restrict__ p, size_t n2, double* out){
  size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
  double s0=0,s1=0,s2=0,s3=0;
  size_t n4 = n2/ (4*stride) * (4*stride);
  for(; i<n4; i+=4*stride){
    double2 a=__ldg(p+i), b=__ldg(p+i+stride), c=__ldg(p+i+2*stride), d=__ldg(p+i+3*stride);
    s0+=a.x+a.y; s1+=b.x+b.y; s2+=c.x+c.y; s3+=d.x+d.y;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s0+s1+s2+s3;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p){return (uint32_t)__cvta_generic_to_shared(p);}
__device__ __forceinline__ void mbar_init(uint64_t* b,int c){asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(smem_u32(b)),"r"(c));}
__device__ __forceinline__ void mbar_expect(uint64_t* b,uint32_t bytes){asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(smem_u32(b)),"r"(bytes):"memory");}
__device__ __forceinline__ void mbar_wait(uint64_t* b,uint32_t ph){
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}"::"r"(smem_u32(b)),"r"(ph):"memory");}
__device__ __forceinline__ void bulk_g2s(void* dst,const void* src,uint32_t bytes,uint64_t* b){
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(smem_u32(dst)),"l"(src),"r"(bytes),"r"(smem_u32(b)):"memory");}

// persistent CTAs; each streams contiguous CHUNK-byte pieces through a STAGES-deep smem ring with cp.async.bulk;
// consumers sum the doubles (1 DADD per 8 bytes).
template<int STAGES,int CHUNK>
__global__ void __launch_bounds__(256) k_read_bulk(const char* __restrict__ p, size_t nchunks, double* out){
  extern __shared__ __align__(128) char sm[];
  __shared__ uint64_t full[STAGES];
  if(threadIdx.x==0){ for(int s=0;s<STAGES;s++) mbar_init(&full[s],1); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  size_t first=blockIdx.x, step=gridDim.x;
  size_t mine = (nchunks>first)? (nchunks-first+step-1)/step : 0;
  if(threadIdx.x==0){
    for(int s=0;s<STAGES && (size_t)s<mine;s++){ mbar_expect(&full[s],CHUNK); bulk_g2s(sm+s*CHUNK,p+(first+s*step)*CHUNK,CHUNK,&full[s]); }
  }
  double acc=0;
  for(size_t t=0;t<mine;t++){
    int s=t%STAGES; uint32_t ph=(t/STAGES)&1;
    mbar_wait(&full[s],ph);
    const double2* q=(const double2*)(sm+s*CHUNK);
    #pragma unroll
    for(int i=threadIdx.x;i<CHUNK/16;i+=256){ double2 v=q[i]; acc+=v.x+v.y; }
    __syncthreads();
    if(threadIdx.x==0 && t+STAGES<mine){ mbar_expect(&full[s],CHUNK); bulk_g2s(sm+s*CHUNK,p+(first+(t+STAGES)*step)*CHUNK,CHUNK,&full[s]); }
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s sms=%d clock=%d kHz\n",pr.name,pr.multiProcessorCount,pr.clockRate);
  int sms=pr.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,sizeof(double)*sms*8*256));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters=20000;
  #define RUN(name,kern,flops_per_thread_iter,blocks_per_sm) { \
    kern<<<sms*blocks_per_sm,256>>>(out,200,1.0000001,1e-9); CK(cudaDeviceSynchronize()); \
    cudaEventRecord(e0); kern<<<sms*blocks_per_sm,256>>>(out,iters,1.0000001,1e-9); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); \
    cudaEventElapsedTime(&ms,e0,e1); \
    double fl=(double)sms*blocks_per_sm*256*(double)iters*(flops_per_thread_iter); \
    printf("%-28s bps=%d  %.3f ms  %.2f TFLOP/s\n",name,blocks_per_sm,ms,fl/ms*1e-9); }
  RUN("dfma ilp8",   k_dfma<8>,  8*2.0, 2);
  RUN("dfma ilp8",   k_dfma<8>,  8*2.0, 4);
  RUN("dfma ilp16",  k_dfma<16>, 16*2.0, 2);
  // m8n8k4: 8*8*4*2 flops per warp-instr = 512 -> 16 per thread
  RUN("dmma m8n8k4 ilp8",  k_dmma884<8>, 8*16.0, 2);
  RUN("dmma m8n8k4 ilp8",  k_dmma884<8>, 8*16.0, 4);
  RUN("dmma m8n8k4 ilp16", k_dmma884<16>, 16*16.0, 2);
  // m16n8k4: 16*8*4*2=1024 per warp -> 32/thread
  RUN("dmma m16n8k4 ilp8", (k_dmma16<8,4>), 8*32.0, 2);
  RUN("dmma m16n8k8 ilp8", (k_dmma16<8,8>), 8*64.0, 2);
  RUN("dmma m16n8k16 ilp8", (k_dmma16<8,16>), 8*128.0, 2);
  RUN("dmma m16n8k16 ilp4", (k_dmma16<4,16>), 4*128.0, 4);

  // bandwidth
  size_t bytes=(size_t)8<<30; char* buf; CK(cudaMalloc(&buf,bytes)); CK(cudaMemset(buf,0,bytes));
  for(int bps=2;bps<=8;bps*=2){
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0); k_read_ldg<<<sms*bps,256>>>((const double2*)buf,bytes/16,out); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms,e0,e1);
    }
    printf("read ldg.128 x4 bps=%d: %.3f ms  %.1f GB/s\n",bps,ms,bytes/ms*1e-6);
  }
  #define RUNB(ST,CH,bps) { \
    CK(cudaFuncSetAttribute(k_read_bulk<ST,CH>,cudaFuncAttributeMaxDynamicSharedMemorySize,ST*CH)); \
    for(int rep=0;rep<3;rep++){ cudaEventRecord(e0); k_read_bulk<ST,CH><<<sms*bps,256,ST*CH>>>(buf,bytes/CH,out); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1);} \
    printf("read bulk stages=%d chunk=%d bps=%d: %.3f ms  %.1f GB/s\n",ST,CH,bps,ms,bytes/ms*1e-6); }
  RUNB(4,16384,1); RUNB(4,16384,2); RUNB(6,32768,1); RUNB(3,32768,2); RUNB(8,16384,1); RUNB(12,16384,1); RUNB(4,8192,4);
  // copy for reference
  char* buf2; CK(cudaMalloc(&buf2,bytes));
  for(int rep=0;rep<3;rep++){ cudaEventRecord(e0); CK(cudaMemcpyAsync(buf2,buf,bytes,cudaMemcpyDeviceToDevice)); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1);} 
  printf("memcpy d2d: %.3f ms  %.1f GB/s (read+write)\n",ms,2.0*bytes/ms*1e-6);
  return 0;
}
