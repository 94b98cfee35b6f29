#include <cstdio>
#include <map>
#define GB_NO_CLASS_LOOKUP 1
#include "gb_predecode.h"
int main(){
  pd_desc_t t[512]; pd_build_base(t);
  std::map<std::pair<unsigned,unsigned>, int> m;
  for(int i=0;i<512;i++){ unsigned h=t[i].x&0xFF, f=t[i].w&0xFFF0u; if(h>=H_RARE) continue; m[{h,f}]++; }
  printf("// gb_classes.inc -- GENERATED (tools/gen_classes.sh): every (handler, operand flags) pair the per-opcode base table of\n"
         "// gb_predecode.h produces for the fast set, one dense id each.  GB_CLS(id, handler, flags): the single-lane build of the\n"
         "// interpreter dispatches on the id to an instance of the instruction body in which both are compile-time constants.\n");
  // ids in order of expected frequency (traces of the benchmark ROMs and the usual SM83 instruction mix), hottest first: nvcc
  // lays the class bodies out in id order, so the bodies a game actually runs sit next to each other in the instruction cache
  static const unsigned prio[][2] = {{4,0x2000},{0,0x0800},{6,0x1000},{2,0x4020},{2,0x4000},{0,0x08a0},{11,0x0000},{0,0x0a00},{3,0x0000},
      {0,0x0820},{2,0x4010},{0,0x0810},{1,0x0020},{0,0x0a80},{3,0x0010},{1,0x0000},{1,0x0200},{12,0x0000},{7,0x0000},{8,0x0160},{9,0x0000},
      {10,0x0160},{3,0x0020},{12,0x0020},{5,0x0000}};
  int id=0;
  for (auto &pr : prio) {
    auto it = m.find({pr[0], pr[1]});
    if (it == m.end()) { fprintf(stderr, "priority entry h=%u f=%04x is not a class\n", pr[0], pr[1]); return 1; }
    printf("GB_CLS(%d, %u, 0x%04xu)\n", id++, pr[0], pr[1]);
    m.erase(it);
  }
  for(auto&kv:m) printf("GB_CLS(%d, %u, 0x%04xu)\n", id++, kv.first.first, kv.first.second);
  printf("#define GB_CLS_COUNT %d\n", id);
}
