// Do the DFMA (sm__pipe_fp64) and DMMA (tensor subpipe dmma) pipes of B200 run concurrently?
// 8 warps per CTA: `ndmma` warps loop on DMMA.8x8x4, the others on DFMA.  Compare against each alone.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__global__ void __launch_bounds__(256) k_mix(double* out,int it_dmma,int it_dfma,int ndmma,int ndfma,double a,double b){
  const int w=threadIdx.x>>5;
  double s=0;
  if(w<ndmma){
    double c[8][2];
    #pragma unroll
    for(int i=0;i<8;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
    for(int it=0;it<it_dmma;++it){
      #pragma unroll
      for(int i=0;i<8;i++) dmma884(c[i][0],c[i][1],a,b);
    }
    #pragma unroll
    for(int i=0;i<8;i++) s+=c[i][0]+c[i][1];
  } else if(w<ndmma+ndfma){
    double acc[16];
    #pragma unroll
    for(int i=0;i<16;i++) acc[i]=threadIdx.x*1e-3+i;
    for(int it=0;it<it_dfma;++it){
      #pragma unroll
      for(int i=0;i<16;i++) acc[i]=fma(acc[i],a,b);
    }
    #pragma unroll
    for(int i=0;i<16;i++) s+=acc[i];
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// same warp interleaves DMMA and DFMA (as the hybrid contraction would)
template<int NF>
__global__ void __launch_bounds__(256) k_interleave(double* out,int iters,double a,double b){
  double c[8][2]; double acc[NF>0?NF:1];
  #pragma unroll
  for(int i=0;i<8;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  #pragma unroll
  for(int i=0;i<NF;i++) acc[i]=threadIdx.x*1e-3+i;
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<8;i++){ dmma884(c[i][0],c[i][1],a,b); if(i*NF/8 != (i+1)*NF/8 || NF>=8) {
      #pragma unroll
      for(int j=i*NF/8;j<(i+1)*NF/8;j++) acc[j]=fma(acc[j],a,b); } }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<8;i++) s+=c[i][0]+c[i][1];
  #pragma unroll
  for(int i=0;i<NF;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr,0); int sms=pr.multiProcessorCount;
  double* out; cudaMalloc(&out,sizeof(double)*sms*4*256);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  auto run=[&](const char* name,int itm,int itf,int nm,int nf,int bps){
    k_mix<<<sms*bps,256>>>(out,100,100,nm,nf,1.0000001,1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_mix<<<sms*bps,256>>>(out,itm,itf,nm,nf,1.0000001,1e-9); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms,e0,e1);
    double fl_m=(double)sms*bps*nm*32*(double)itm*8*16.0, fl_f=(double)sms*bps*nf*32*(double)itf*16*2.0;
    printf("%-34s %.3f ms  dmma %.2f TF  dfma %.2f TF  total %.2f TF\n",name,ms,fl_m/ms*1e-9,fl_f/ms*1e-9,(fl_m+fl_f)/ms*1e-9);
  };
  run("dmma only (8 warps)",20000,0,8,0,2);
  run("dfma only (8 warps)",0,80000,0,8,2);
  run("dmma 4 warps only",20000,0,4,0,2);
  run("dfma 4 warps only",0,80000,0,4,2);
  run("4 dmma + 4 dfma (balanced time)",20000,72000,4,4,2);
  run("4 dmma + 4 dfma (dfma light 1/4)",20000,18000,4,4,2);
  run("6 dmma + 2 dfma",20000,72000,6,2,2);
  #define RUNI(NF) { k_interleave<NF><<<sms*2,256>>>(out,100,1.0000001,1e-9); cudaDeviceSynchronize(); \
    cudaEventRecord(e0); k_interleave<NF><<<sms*2,256>>>(out,20000,1.0000001,1e-9); cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms,e0,e1); \
    double fm=(double)sms*2*256*20000.0*8*16.0, ff=(double)sms*2*256*20000.0*NF*2.0; \
    printf("interleave 8 dmma + %2d dfma/iter     %.3f ms  dmma %.2f TF  dfma %.2f TF  total %.2f TF\n",NF,ms,fm/ms*1e-9,ff/ms*1e-9,(fm+ff)/ms*1e-9); }
  RUNI(0); RUNI(8); RUNI(16); RUNI(32);
  return 0;
}
