#define PASS_SH true
#define PASS_FZ false
#include "pass_inst.inc"
