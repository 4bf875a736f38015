// Debug aid: prints a native backtrace on SIGSEGV (module + offset per frame; resolve with addr2line here).
//   gcc -shared -fPIC -o tools/bin/libsegv_trace.so tools/segv_trace.c ;  ctypes.CDLL(...) installs the handler
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

static void on_segv(int sig) {
    void *buf[96];
    int n = backtrace(buf, 96);
    backtrace_symbols_fd(buf, n, 2);
    _exit(128 + sig);
}

__attribute__((constructor)) static void install(void) {
    struct sigaction sa;
    sa.sa_handler = on_segv;
    sigemptyset(&sa.sa_mask);
    sa.sa_flags = SA_RESETHAND;
    sigaction(SIGSEGV, &sa, 0);
    sigaction(SIGBUS, &sa, 0);
    sigaction(SIGABRT, &sa, 0);
}
