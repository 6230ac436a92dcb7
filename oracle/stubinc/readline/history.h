/* TEST INFRASTRUCTURE (oracle build only). */
#ifndef ODX_STUB_HISTORY_H
#define ODX_STUB_HISTORY_H
void add_history(const char *line);
void using_history(void);
int read_history(const char *filename);
int write_history(const char *filename);
void stifle_history(int max);
#endif
