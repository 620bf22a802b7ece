#include "so101_oracle.h"
void so_collide(const so_model *m, so_data *d) { (void)m; d->ncon = 0; }
