// Micro-benchmark: bandwidth of fragment-shaped copies vs plain row copies on padded 1080p planes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
constexpr int W=1920,H=1088,STRIDE=1952,NF=(W/8)*(H/8);
// (1) 4 lanes per fragment, lane l copies rows 2l,2l+1 (8 B each); fragments in raster order
__global__ void k_frag4(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t plane, int nframes){
  int t=blockIdx.x*blockDim.x+threadIdx.x; int f=t>>2, l=t&3; int fr=blockIdx.y;
  if(f>=NF) return; int fx=f%(W/8), fy=f/(W/8);
  size_t off=(size_t)fr*plane+(size_t)(fy*8+2*l)*STRIDE+fx*8+16+16*STRIDE;
  uint2 a=*(const uint2*)(src+off), b=*(const uint2*)(src+off+STRIDE);
  *(uint2*)(dst+off)=a; *(uint2*)(dst+off+STRIDE)=b;
}
// (2) same but with an MV-like unaligned source (+3 bytes, +1 row): two aligned 8B loads per row
__global__ void k_frag4_unaligned(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t plane, int nframes){
  int t=blockIdx.x*blockDim.x+threadIdx.x; int f=t>>2, l=t&3; int fr=blockIdx.y;
  if(f>=NF) return; int fx=f%(W/8), fy=f/(W/8);
  size_t off=(size_t)fr*plane+(size_t)(fy*8+2*l)*STRIDE+fx*8+16+16*STRIDE;
  const uint8_t* s=src+off+STRIDE+3;
  uint2 o[2];
  #pragma unroll
  for(int r=0;r<2;r++){ const uint8_t* p=s+r*STRIDE; const uint2* w=(const uint2*)((uintptr_t)p&~(uintptr_t)7); unsigned sh=(uintptr_t)p&7;
    uint2 w0=w[0], w1=w[1]; unsigned sel=0x3210u+0x1111u*(sh&3);
    if(sh<4){o[r].x=__byte_perm(w0.x,w0.y,sel); o[r].y=__byte_perm(w0.y,w1.x,sel);} else {o[r].x=__byte_perm(w0.y,w1.x,sel); o[r].y=__byte_perm(w1.x,w1.y,sel);} }
  *(uint2*)(dst+off)=o[0]; *(uint2*)(dst+off+STRIDE)=o[1];
}
// (3) plain row copy, 16 B per thread
__global__ void k_rows16(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t plane, int nframes){
  int x=(blockIdx.x*blockDim.x+threadIdx.x)*16; int y=blockIdx.y%H; int fr=blockIdx.y/H;
  if(x>=W) return; size_t off=(size_t)fr*plane+(size_t)(y+16)*STRIDE+x+16;
  *(uint4*)(dst+off)=*(const uint4*)(src+off);
}
// (4) one thread per fragment row pair but 8 lanes per fragment (lane = row)
__global__ void k_frag8(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t plane, int nframes){
  int t=blockIdx.x*blockDim.x+threadIdx.x; int f=t>>3, l=t&7; int fr=blockIdx.y;
  if(f>=NF) return; int fx=f%(W/8), fy=f/(W/8);
  size_t off=(size_t)fr*plane+(size_t)(fy*8+l)*STRIDE+fx*8+16+16*STRIDE;
  *(uint2*)(dst+off)=*(const uint2*)(src+off);
}
// (5) 2 lanes per fragment-pair: each lane copies 16 B (two horizontally adjacent fragments) x 4 rows
__global__ void k_pair16(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t plane, int nframes){
  int t=blockIdx.x*blockDim.x+threadIdx.x; int pf=t>>1, l=t&1; int fr=blockIdx.y;
  if(pf>=NF/2) return; int px=pf%(W/16), fy=pf/(W/16);
  size_t off=(size_t)fr*plane+(size_t)(fy*8+4*l)*STRIDE+px*16+16+16*STRIDE;
  uint4 v[4];
  #pragma unroll
  for(int r=0;r<4;r++) v[r]=*(const uint4*)(src+off+r*STRIDE);
  #pragma unroll
  for(int r=0;r<4;r++) *(uint4*)(dst+off+r*STRIDE)=v[r];
}
// (6) record-driven: each 4-lane group first loads a 16-byte record (offset, mv) and then copies from the
//     motion-displaced unaligned source: the dependent-load structure of a lean recon kernel without barriers
__global__ void k_rec_driven(const int4* __restrict__ recs, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t plane){
  int t=blockIdx.x*blockDim.x+threadIdx.x; int f=t>>2, l=t&3; int fr=blockIdx.y;
  if(f>=NF) return;
  int4 rw=__ldg(recs+(size_t)fr*NF+f);
  int mv=rw.y<<16>>16; int dx=(int)(signed char)(mv&0xFF), dy=mv>>8;
  int ax=abs(dx), ay=abs(dy); int sx=dx<0?-1:1, sy=dy<0?-1:1; int mx=sx*(ax>>1), my=sy*(ay>>1);
  size_t off=(size_t)fr*plane+(size_t)rw.x+(size_t)(2*l)*STRIDE;
  const uint8_t* s=src+off+my*STRIDE+mx;
  uint2 o[2];
  #pragma unroll
  for(int r=0;r<2;r++){ const uint8_t* p=s+r*STRIDE; const uint2* w=(const uint2*)((uintptr_t)p&~(uintptr_t)7); unsigned sh=(uintptr_t)p&7;
    uint2 w0=w[0]; if(sh==0){o[r]=w0; continue;} uint2 w1=w[1]; unsigned sel=0x3210u+0x1111u*(sh&3);
    if(sh<4){o[r].x=__byte_perm(w0.x,w0.y,sel); o[r].y=__byte_perm(w0.y,w1.x,sel);} else {o[r].x=__byte_perm(w0.y,w1.x,sel); o[r].y=__byte_perm(w1.x,w1.y,sel);} }
  *(uint2*)(dst+off)=o[0]; *(uint2*)(dst+off+STRIDE)=o[1];
}
int main(){
  const int NFR=32; size_t plane=(size_t)STRIDE*(H+32); size_t bytes=plane*NFR;
  uint8_t *a,*b; CK(cudaMalloc(&a,bytes)); CK(cudaMalloc(&b,bytes)); CK(cudaMemset(a,1,bytes)); CK(cudaMemset(b,2,bytes));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double moved=(double)W*H*NFR*2;
  auto run=[&](const char* name, auto launch){ for(int i=0;i<3;i++) launch(); cudaEventRecord(e0); for(int i=0;i<20;i++) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=20; printf("%-22s %8.2f us  %7.1f GB/s\n",name,ms*1e3,moved/ms/1e6); };
  run("frag4 aligned", [&]{ k_frag4<<<dim3((NF*4+255)/256,NFR),256>>>(a,b,plane,NFR); });
  run("frag4 unaligned", [&]{ k_frag4_unaligned<<<dim3((NF*4+255)/256,NFR),256>>>(a,b,plane,NFR); });
  run("frag8 (lane=row)", [&]{ k_frag8<<<dim3((NF*8+255)/256,NFR),256>>>(a,b,plane,NFR); });
  run("pair16 (16B x4 rows)", [&]{ k_pair16<<<dim3((NF+255)/256,NFR),256>>>(a,b,plane,NFR); });
  run("rows 16B", [&]{ k_rows16<<<dim3((W/16+127)/128,H*NFR),128>>>(a,b,plane,NFR); });
  { // records: raster offsets, mv=(6,2) half-pel
    int4* h=(int4*)malloc(sizeof(int4)*NF*NFR); for(int fr=0;fr<NFR;fr++) for(int f=0;f<NF;f++){ int fx=f%(W/8), fy=f/(W/8); int4 r; r.x=(fy*8+16)*STRIDE+fx*8+16; r.y=((2&0xFF)<<8)|6; r.z=0; r.w=0; h[(size_t)fr*NF+f]=r; }
    int4* d; CK(cudaMalloc(&d,sizeof(int4)*NF*NFR)); CK(cudaMemcpy(d,h,sizeof(int4)*NF*NFR,cudaMemcpyHostToDevice));
    moved=(double)W*H*NFR*2+16.0*NF*NFR;
    run("rec-driven unaligned", [&]{ k_rec_driven<<<dim3((NF*4+255)/256,NFR),256>>>(d,a,b,plane); });
  }
  CK(cudaDeviceSynchronize()); return 0;
}
