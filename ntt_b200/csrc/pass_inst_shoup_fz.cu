#define PASS_SH true
#define PASS_FZ true
#include "pass_inst.inc"
