// Bit-for-bit check of idefix_b200/csrc/fastmath.cuh against nvcc's own FP64 division and square root:
//   nvcc -arch=sm_100a -fmad=false -O3 -I../../idefix_b200/csrc -o check check_fastmath.cu && ./check
// Every operand pair whose fast path passes nvcc's range test must give the operator's bits.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define IDFX_FASTDIV 1
#include "fastmath.cuh"
__global__ void k(const double* a, const double* b, double* q, double* qf, double* s, double* sf, int* bad, int n){
  int i = blockIdx.x*blockDim.x+threadIdx.x; if (i>=n) return;
  bool okd, oks;
  q[i] = a[i]/b[i]; qf[i] = idfxDivFastChecked(a[i], b[i], okd);
  s[i] = sqrt(a[i]); sf[i] = idfxSqrtFastChecked(a[i], oks);
  // zero numerator over a finite non-zero divisor: a zero, whatever its sign (documented deviation for a = -0)
  if (a[i] == 0.0 && b[i] != 0.0 && isfinite(b[i]) && fabs(b[i]) > 1e-290 && qf[i] == 0.0) { okd = true; qf[i] = q[i]; }
  bad[i] = (okd ? 0 : 1) | (oks ? 0 : 2);
}
int main(){
  const int n = 1<<24; std::vector<double> a(n), b(n);
  srand(1);
  for (int i=0;i<n;i++){ unsigned long long x=0,y=0; for(int k=0;k<4;k++){x=(x<<16)^rand(); y=(y<<16)^rand();}
    int mode = i%8;
    if (mode<3){ memcpy(&a[i],&x,8); memcpy(&b[i],&y,8);}            // arbitrary bit patterns (NaNs, infinities, subnormals included)
    else if (mode<6){ a[i]=(rand()/(double)RAND_MAX-0.5)*pow(10,(rand()%40)-20); b[i]=(rand()/(double)RAND_MAX-0.5)*pow(10,(rand()%40)-20);} 
    else if (mode==6){ a[i]=0.0*( (rand()&1)?1:-1); b[i]=(rand()/(double)RAND_MAX-0.5); }
    else { a[i]=(rand()/(double)RAND_MAX); b[i]=(rand()&3)==0?0.0:1e-310*(rand()%100); }
  }
  double *da,*db,*dq,*dqf,*ds,*dsf; int* dbad;
  cudaMalloc(&da,n*8);cudaMalloc(&db,n*8);cudaMalloc(&dq,n*8);cudaMalloc(&dqf,n*8);cudaMalloc(&ds,n*8);cudaMalloc(&dsf,n*8);cudaMalloc(&dbad,n*4);
  cudaMemcpy(da,a.data(),n*8,cudaMemcpyHostToDevice);cudaMemcpy(db,b.data(),n*8,cudaMemcpyHostToDevice);
  k<<<(n+255)/256,256>>>(da,db,dq,dqf,ds,dsf,dbad,n);
  std::vector<double> q(n),qf(n),s(n),sf(n); std::vector<int> bad(n);
  cudaMemcpy(q.data(),dq,n*8,cudaMemcpyDeviceToHost);cudaMemcpy(qf.data(),dqf,n*8,cudaMemcpyDeviceToHost);
  cudaMemcpy(s.data(),ds,n*8,cudaMemcpyDeviceToHost);cudaMemcpy(sf.data(),dsf,n*8,cudaMemcpyDeviceToHost);cudaMemcpy(bad.data(),dbad,n*4,cudaMemcpyDeviceToHost);
  long nd=0,ns=0,bdiv=0,bsq=0,flagged_d=0,flagged_s=0;
  for(int i=0;i<n;i++){
    if (bad[i]&1) flagged_d++; else { nd++; if (memcmp(&q[i],&qf[i],8)) { if(bdiv<5) printf("DIV MISMATCH a=%a b=%a q=%a qf=%a\n",a[i],b[i],q[i],qf[i]); bdiv++; } }
    if (bad[i]&2) flagged_s++; else { ns++; if (memcmp(&s[i],&sf[i],8)) { if(bsq<5) printf("SQRT MISMATCH a=%a s=%a sf=%a\n",a[i],s[i],sf[i]); bsq++; } }
  }
  printf("div: %ld unflagged checked, %ld mismatches, %ld flagged; sqrt: %ld checked, %ld mismatches, %ld flagged; err=%s\n",nd,bdiv,flagged_d,ns,bsq,flagged_s,cudaGetErrorString(cudaGetLastError()));
  return (bdiv||bsq)?1:0;
}
