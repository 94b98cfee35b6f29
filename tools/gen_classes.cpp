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
  int id=0;
  for(auto&kv:m) printf("GB_CLS(%d, %u, 0x%04xu)\n", id++, kv.first.first, kv.first.second);
  printf("#define GB_CLS_COUNT %d\n", id);
}
