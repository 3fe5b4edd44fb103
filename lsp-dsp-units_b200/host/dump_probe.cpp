/*
 * dump_probe -- runs lsp::dspu::Convolver::dump() of the drop-in facade against the reference's own
 * IStateDumper interface and prints "name=value" per field.  Built by tests/test_host_facade.py with
 * -DB200CONV_WITH_STATE_DUMPER against /root/reference/include (test infrastructure; needs no GPU
 * as long as the convolver stays un-initialised; with "init" as argv[1] it loads a 300-tap IR and
 * processes 200 samples first).
 */
#include <lsp-plug.in/dsp-units/iface/IStateDumper.h>
#include <lsp-plug.in/dsp-units/util/Convolver.h>

#include <stdio.h>
#include <string.h>
#include <vector>

namespace
{
    class PrintDumper: public lsp::dspu::IStateDumper
    {
        public:
            virtual void write(const char *name, const void *v) override        { printf("%s=%s\n", name, (v != NULL) ? "ptr" : "null"); }
            virtual void write(const char *name, unsigned long v) override      { printf("%s=%lu\n", name, v); }
            virtual void write(const char *name, unsigned long long v) override { printf("%s=%llu\n", name, v); }
            virtual void write(const char *name, unsigned int v) override       { printf("%s=%u\n", name, v); }
            virtual void write(const char *name, int v) override                { printf("%s=%d\n", name, v); }
            virtual void write(const char *name, float v) override              { printf("%s=%.9g\n", name, v); }
    };
}

int main(int argc, char **argv)
{
    lsp::dspu::Convolver c;
    if ((argc > 1) && (!strcmp(argv[1], "init")))
    {
        std::vector<float> ir(300, 0.01f), x(200, 0.5f), y(200);
        if (!c.init(ir.data(), ir.size(), 9, 0.25f))
            return 2;
        c.process(y.data(), x.data(), x.size());
    }
    PrintDumper d;
    c.dump(&d);
    return 0;
}
