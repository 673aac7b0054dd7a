// Host-only statistics of the numeric schedule for a 3-D 7-point Laplacian g^3 (cfg2: g = 100, cfg5: g = 160) partitioned over
// several world sizes: launches, segments, collective payload, image / scratch / table sizes and, per rank, the tile flops it
// actually executes (in-panel updates and Schur scatters, with the tile shares of split updates).  No GPU, no oracle.
//   R=sparse-matrix-factorization-library_b200/csrc; g++ -O2 -std=c++17 -I$R scripts/plan_stats.cpp $R/plan.cpp $R/symbolic.cpp $R/parallel_order.cpp \
//       $R/xstat_shim.c /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -o /tmp/plan_stats && /tmp/plan_stats 160 2 4 8
#include "plan.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
using namespace sfb200;
static double now(){return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}
int main(int argc,char**argv){
  int g=atoi(argv[1]);
  i64 n=(i64)g*g*g; std::vector<i64> rows,cols; std::vector<double> vals;
  // same triplet order as sparseframe_b200.matrices.laplace3d is not needed here: structure statistics only
  std::vector<i64> cp(n+1,0),ci; std::vector<double> cx;
  for(i64 i=0;i<g;i++)for(i64 j=0;j<g;j++)for(i64 k=0;k<g;k++){i64 v=(i*g+j)*g+k; ci.push_back(v);cx.push_back(6);
    if(k+1<g){ci.push_back(v+1);cx.push_back(-1);} if(j+1<g){ci.push_back(v+g);cx.push_back(-1);} if(i+1<g){ci.push_back(v+g*g);cx.push_back(-1);} cp[v+1]=ci.size();}
  CscView A; A.n=n;A.colptr=cp.data();A.rowind=ci.data();A.val=cx.data();A.symmetric=true;
  Symbolic S; double t0=now(); int rc=analyze(A,Kind::Cholesky,5380000000ull,S); printf("analyze rc %d %.1fs nsuper %ld xsize %.2f GB\n",rc,now()-t0,(long)S.nsuper,S.xsize*8e-9); fflush(stdout);
  std::vector<double> tols={1.05}; if(const char*e=getenv("TOLS")){tols.clear(); char*dup=strdup(e); for(char*t=strtok(dup,",");t;t=strtok(nullptr,",")) tols.push_back(atof(t));}
  for(double tol:tols) for(int a=2;a<argc;a++){ int world=atoi(argv[a]); double maxshare=0,sumshare=0; printf("---- balance_tol %.3f world %d\n",tol,world);
    for(int r=0;r<world;r++){ HostPlan P; PlanParams prm; prm.balance_tol=tol; t0=now(); rc=build_plan(S,prm,P,r,world);
      i64 maxgrid=0; for(auto&L:P.launches) maxgrid=std::max<i64>(maxgrid,L.grid);
      i64 collbytes=0; for(auto&c:P.colls) if(c.type==C_BCAST) collbytes+=(i64)c.rows*c.cols*8;
      double tabGB=(P.upd_coloff.size()*8+P.upd_relbase.size()*8+P.upd_segb.size()*4*3+P.relmap.size()*4+P.asm_dst_l.size()*8+P.gemm_ops.size()*sizeof(GemmOp)+P.lsi32.size()*4+P.nz_col.size()*4+P.nz_row.size()*8)*1e-9;
      double gf=0,sf=0; { for(auto&L:P.launches){ if(L.kind==L_GEMM128||L.kind==L_GEMM64||L.kind==L_GEMM128WS){ for(int q=L.first;q<L.first+L.count;q++){auto&g2=P.gemm_ops[q]; const int B=L.kind==L_GEMM64?64:128; double tiles=(double)share_of_tiles(gemm_tiles(g2.M,g2.N,B,(g2.flags&G_LOWER)!=0),g2.tile_mod,g2.tile_rem); double f=tiles*2.0*B*B*g2.K; if(g2.flags&G_SCATTER) sf+=f; else gf+=f; } } } }
      printf("   executed tile flops: in-panel %.3e scatter %.3e total share %.3f\n",gf,sf,(gf+sf)/P.flops); maxshare=std::max(maxshare,(gf+sf)/P.flops); sumshare+=(gf+sf)/P.flops;
      printf("world %d rank %d: rc %d %.1fs launches %zu segments %zu colls %zu (%.1f GB) maxgrid %ld local flops %.3f image %.2f GB scratch %.2f GB tables %.2f GB\n",world,r,rc,now()-t0,P.launches.size(),P.segments.size(),P.colls.size(),collbytes*1e-9,(long)maxgrid,P.local_flops/P.flops,P.xsize_pad*8e-9,P.wsize*8e-9,tabGB); fflush(stdout);
      }
    printf("==== tol %.3f world %d: max executed share %.4f (ideal %.4f, sum %.3f) -> imbalance %.3f\n",tol,world,maxshare,1.0/world,sumshare,maxshare*world/sumshare);
  }
}
