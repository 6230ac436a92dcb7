/* TEST INFRASTRUCTURE (oracle build only). readline headers are not installed; the oracle runs headless. */
#ifndef ODX_STUB_READLINE_H
#define ODX_STUB_READLINE_H
typedef char **rl_completion_func_t(const char *, int, int);
typedef char *rl_compentry_func_t(const char *, int);
extern const char *rl_library_version;
extern const char *rl_readline_name;
extern int rl_attempted_completion_over;
extern rl_completion_func_t *rl_attempted_completion_function;
char *readline(const char *prompt);
int rl_read_init_file(const char *filename);
char **rl_completion_matches(const char *text, rl_compentry_func_t *entry_func);
#endif
