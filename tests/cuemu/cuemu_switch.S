// Minimal x86-64 SysV context switch for the cuemu fiber scheduler
// (test infrastructure only; see cuemu.h).
//   void cuemu_switch(void** old_sp, void* new_sp);
        .text
        .globl  cuemu_switch
        .type   cuemu_switch, @function
cuemu_switch:
        pushq   %rbp
        pushq   %rbx
        pushq   %r12
        pushq   %r13
        pushq   %r14
        pushq   %r15
        movq    %rsp, (%rdi)
        movq    %rsi, %rsp
        popq    %r15
        popq    %r14
        popq    %r13
        popq    %r12
        popq    %rbx
        popq    %rbp
        ret
        .size   cuemu_switch, .-cuemu_switch
        .section .note.GNU-stack,"",@progbits
