#define PASS_SH false
#define PASS_FZ false
#include "pass_inst.inc"
