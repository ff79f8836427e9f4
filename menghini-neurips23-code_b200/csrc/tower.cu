#include "ctx.h"
struct gb_tower { int dummy; };
void gb_tower_free(gb_tower* t) { delete t; }
