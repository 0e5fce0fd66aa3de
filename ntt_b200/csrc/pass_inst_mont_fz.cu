#define PASS_SH false
#define PASS_FZ true
#include "pass_inst.inc"
