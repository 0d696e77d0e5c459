// tests/cusim/cusim_swap.cpp -- TEST INFRASTRUCTURE ONLY.
// Minimal x86-64 SysV fiber switch for the SIMT emulator: saves the callee-saved
// integer registers on the current stack, publishes the stack pointer, adopts
// the target stack and returns into it.
__asm__(
    ".text\n"
    ".globl cusim_swap\n"
    ".type cusim_swap,@function\n"
    "cusim_swap:\n"
    "  pushq %rbp\n"
    "  pushq %rbx\n"
    "  pushq %r12\n"
    "  pushq %r13\n"
    "  pushq %r14\n"
    "  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n"
    "  popq %r14\n"
    "  popq %r13\n"
    "  popq %r12\n"
    "  popq %rbx\n"
    "  popq %rbp\n"
    "  ret\n"
    ".size cusim_swap,.-cusim_swap\n");
