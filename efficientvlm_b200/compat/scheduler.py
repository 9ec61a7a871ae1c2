"""`scheduler` drop-in (reference: scheduler.py:4-29): `create_scheduler(args, optimizer)` with the same argument handling, returning the
linear warm-up / linear decay schedule for an optimizer that is not a `torch.optim.Optimizer` (`LambdaLR` insists on one)."""
from efficientvlm_b200.optim import LinearWarmupDecay


def create_scheduler(args, optimizer):
    if "num_training_steps" not in args:
        args["num_training_steps"] = args["epochs"] * args["step_per_epoch"]
    print("### num_training_steps, ", args["num_training_steps"], flush=True)
    if isinstance(args["num_warmup_steps"], float):
        assert 0 <= args["num_warmup_steps"] < 1
        args["num_warmup_steps"] = int(args["num_training_steps"] * args["num_warmup_steps"])
    print("### num_warmup_steps, ", args["num_warmup_steps"], flush=True)
    if args["sched"] != "linear":
        raise NotImplementedError(f"args.sched == {args['sched']}")
    return LinearWarmupDecay(optimizer, args["num_training_steps"], args["num_warmup_steps"])
