// Micro-test: TMA variants, one per process (argv[1]) because a fault poisons the context.
//  0: 1-D bulk copy (UBLKCP)   1: 2-D tensor tile, u32 128x8, param map   2: same, u32 32x8
//  3: u8 256x8   4: u32 128x8 map in global memory + tensormap fence   5: as 1 but load only (no store)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ unsigned s32(const void*p){return (unsigned)__cvta_generic_to_shared(p);}
__device__ __forceinline__ void mbar_wait(unsigned long long* mbar){ unsigned done=0; while(!done){ asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}":"=r"(done):"r"(s32(mbar)),"r"(0u):"memory"); } }
__global__ void k_bulk1d(uint8_t* g){
  __shared__ __align__(128) uint8_t buf[4096]; __shared__ __align__(8) unsigned long long mbar;
  if(threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  if(threadIdx.x==0){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&mbar)),"r"(4096u):"memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(s32(buf)),"l"(g),"r"(4096u),"r"(s32(&mbar)):"memory"); }
  mbar_wait(&mbar);
  g[8192+threadIdx.x]=buf[threadIdx.x]+1;
}
template<int BYTES,bool GLOBAL_MAP,bool STORE>
__device__ void body(const CUtensorMap* tm, int tx, int ty){
  __shared__ __align__(128) uint8_t tile[BYTES];
  __shared__ __align__(8) unsigned long long mbar;
  int tid=threadIdx.x;
  if(tid==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  if(tid==0){
    if(GLOBAL_MAP) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;"::"l"(tm):"memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&mbar)),"r"((unsigned)BYTES):"memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"::"r"(s32(tile)),"l"(tm),"r"(tx),"r"(ty),"r"(s32(&mbar)):"memory");
  }
  mbar_wait(&mbar);
  for(int i=tid;i<BYTES;i+=blockDim.x) tile[i]+=1;
  if(STORE){
    asm volatile("fence.proxy.async.shared::cta;":::"memory");
    __syncthreads();
    if(tid==0){ asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"::"l"(tm),"r"(tx),"r"(ty),"r"(s32(tile)):"memory"); asm volatile("cp.async.bulk.commit_group;":::"memory"); asm volatile("cp.async.bulk.wait_group.read 0;":::"memory"); }
  }
}
__global__ void k1(const __grid_constant__ CUtensorMap tm,int tx,int ty){ body<4096,false,true>(&tm,tx,ty); }
__global__ void k2(const __grid_constant__ CUtensorMap tm,int tx,int ty){ body<1024,false,true>(&tm,tx,ty); }
__global__ void k3(const __grid_constant__ CUtensorMap tm,int tx,int ty){ body<2048,false,true>(&tm,tx,ty); }
__global__ void k4(const CUtensorMap* tm,int tx,int ty){ body<4096,true,true>(tm,tx,ty); }
__global__ void k5(const __grid_constant__ CUtensorMap tm,int tx,int ty){ body<4096,false,false>(&tm,tx,ty); }
typedef CUresult (*Enc)(CUtensorMap*,CUtensorMapDataType,cuuint32_t,void*,const cuuint64_t*,const cuuint64_t*,const cuuint32_t*,const cuuint32_t*,CUtensorMapInterleave,CUtensorMapSwizzle,CUtensorMapL2promotion,CUtensorMapFloatOOBfill);
int main(int argc,char**argv){
  int v=argc>1?atoi(argv[1]):0;
  const int PITCH=1952, ROWS=64; uint8_t* d; CK(cudaMalloc(&d,(size_t)PITCH*ROWS)); CK(cudaMemset(d,0,(size_t)PITCH*ROWS));
  void* fn=nullptr; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fn,cudaEnableDefault,&q));
  CUtensorMap tm; cuuint64_t str[1]={PITCH}; cuuint32_t es[2]={1,1}; CUresult r=CUDA_SUCCESS;
  if(v==3){ cuuint64_t dims[2]={PITCH,ROWS}; cuuint32_t box[2]={256,8}; r=((Enc)fn)(&tm,CU_TENSOR_MAP_DATA_TYPE_UINT8,2,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_NONE,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  else { cuuint64_t dims[2]={PITCH/4,ROWS}; cuuint32_t box[2]={(cuuint32_t)(v==2?32:128),8}; r=((Enc)fn)(&tm,CU_TENSOR_MAP_DATA_TYPE_UINT32,2,d,dims,str,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_NONE,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  if(r!=CUDA_SUCCESS){ printf("variant %d: encode failed %d\n",v,(int)r); return 0; }
  if(v==0) k_bulk1d<<<1,64>>>(d);
  else if(v==1) k1<<<1,64>>>(tm,argc>2?atoi(argv[2]):4,12);
  else if(v==2) k2<<<1,64>>>(tm,4,12);
  else if(v==3) k3<<<1,64>>>(tm,16,12);
  else if(v==4){ CUtensorMap* dm; CK(cudaMalloc(&dm,sizeof(tm))); CK(cudaMemcpy(dm,&tm,sizeof(tm),cudaMemcpyHostToDevice)); k4<<<1,64>>>(dm,4,12); }
  else k5<<<1,64>>>(tm,4,12);
  cudaError_t e=cudaDeviceSynchronize(); printf("variant %d: %s\n",v,cudaGetErrorString(e));
  if(e==cudaSuccess){ std::vector<uint8_t> h((size_t)PITCH*ROWS); cudaMemcpy(h.data(),d,h.size(),cudaMemcpyDeviceToHost); long s=0; for(auto x:h) s+=x; printf("  byte sum = %ld\n",s); }
  return 0;
}
